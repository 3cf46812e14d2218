// TEST INFRASTRUCTURE: runs the per-thread bodies of sniper_b200/csrc/depthwise_core.cuh on the CPU, thread index by
// thread index and block by block, exactly as depthwise.cu launches them (same thread counts, same wgrad grid rule).
// Built on the fly by tests/test_depthwise_cpu.py (nvcc compiles the __host__ side of the __host__ __device__ bodies);
// never linked into the product library.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdlib.h>

#include "../sniper_b200/csrc/depthwise_core.cuh"

typedef __nv_bfloat16 bf16;

static void fill(dwc::Params& p, int NB, int H, int W, int C, int stride, long ld_in, long ld_out) {
  p.NB = NB; p.H = H; p.W = W; p.C = C; p.stride = stride;
  p.Ho = (H - 1) / stride + 1; p.Wo = (W - 1) / stride + 1;
  p.ldx = ld_in; p.ldy = ld_out;
}

template <typename T>
static void fwd_t(const void* x, long ldx, const float* w, void* y, long ldy, int NB, int H, int W, int C, int stride) {
  dwc::Params p; fill(p, NB, H, W, C, stride, ldx, ldy);
  if (stride == 1) { const long nt = dwc::fwd_threads<1, 4>(p); for (long t = 0; t < nt; ++t) dwc::fwd<T, 1, 4, false>(t, (const T*)x, w, (T*)y, p); }
  else             { const long nt = dwc::fwd_threads<2, 2>(p); for (long t = 0; t < nt; ++t) dwc::fwd<T, 2, 2, false>(t, (const T*)x, w, (T*)y, p); }
}
template <typename T>
static void dgrad_t(const void* dy, long lddy, const float* w, void* dx, long lddx, int NB, int H, int W, int C, int stride) {
  dwc::Params p;
  if (stride == 1) {           // as sniper_depthwise3x3_dgrad: the forward body on dy with the flipped filter
    fill(p, NB, H, W, C, 1, lddy, lddx);
    const long nt = dwc::fwd_threads<1, 4>(p);
    for (long t = 0; t < nt; ++t) dwc::fwd<T, 1, 4, true>(t, (const T*)dy, w, (T*)dx, p);
  } else {
    fill(p, NB, H, W, C, stride, lddx, lddy);
    const long nt = dwc::dgrad_s2_threads(p);
    for (long t = 0; t < nt; ++t) dwc::dgrad_s2<T>(t, (const T*)dy, w, (T*)dx, p);
  }
}
template <typename T>
static void wgrad_t(const void* x, long ldx, const void* dy, long lddy, float* dw, int NB, int H, int W, int C, int stride, int gx) {
  dwc::Params p; fill(p, NB, H, W, C, stride, ldx, lddy);
  const int TY = 8;
  const int LC = dwc::wgrad_lc(C);
  const int gy = ((C >> 2) + LC - 1) / LC;
  for (int by = 0; by < gy; ++by)
    for (int bx = 0; bx < gx; ++bx)
      for (int ty = 0; ty < TY; ++ty)
        for (int tx = 0; tx < 32; ++tx) {
          float acc[9][4];
          const int c = stride == 1 ? dwc::wgrad_partial<T, 1, 4>(bx, by, tx, ty, TY, gx, (const T*)x, (const T*)dy, p, acc)
                                    : dwc::wgrad_partial<T, 2, 2>(bx, by, tx, ty, TY, gx, (const T*)x, (const T*)dy, p, acc);
          if (c < 0) continue;
          for (int t = 0; t < 9; ++t) for (int k = 0; k < 4; ++k) dw[(long)t * C + c + k] += acc[t][k];
        }
}

template <int S, bool FLIP>
static void tiled_t(const void* x, long ldx, const float* w, void* y, long ldy, int NB, int H, int W, int C) {
  dwc::Params p; fill(p, NB, H, W, C, S, ldx, ldy);
  const long blocks = dwc::tile_blocks<S>(p);
  bf16* smem = (bf16*)malloc(sizeof(bf16) * dwc::Tile<S>::ELEMS);
  for (int by = 0; by < C / 64; ++by)
    for (long bx = 0; bx < blocks; ++bx) {
      const dwc::TileId t = dwc::tile_id<S>(bx, by, p);
      memset(smem, 0x7f, sizeof(bf16) * dwc::Tile<S>::ELEMS);        // poison: every element must be staged
      for (int tid = 0; tid < 256; ++tid) dwc::tile_stage<S>(tid, 256, t, (const bf16*)x, smem, p);
      for (int tid = 0; tid < 256; ++tid) dwc::tile_compute<S, FLIP>(tid, 256, t, smem, w, (bf16*)y, p);
    }
  free(smem);
}

extern "C" {
// the shared-memory tiled bf16 kernels, block by block (stage phase, then compute phase); flip = stride-1 data gradient
void emu_dw_tiled(const void* x, long ldx, const float* w, void* y, long ldy, int NB, int H, int W, int C, int stride, int flip) {
  if (stride == 1) { if (flip) tiled_t<1, true>(x, ldx, w, y, ldy, NB, H, W, C); else tiled_t<1, false>(x, ldx, w, y, ldy, NB, H, W, C); }
  else tiled_t<2, false>(x, ldx, w, y, ldy, NB, H, W, C);
}
void emu_dw_fwd(const void* x, long ldx, const float* w, void* y, long ldy, int NB, int H, int W, int C, int stride, int dtype) {
  if (dtype == 0) fwd_t<float>(x, ldx, w, y, ldy, NB, H, W, C, stride); else fwd_t<bf16>(x, ldx, w, y, ldy, NB, H, W, C, stride);
}
void emu_dw_dgrad(const void* dy, long lddy, const float* w, void* dx, long lddx, int NB, int H, int W, int C, int stride, int dtype) {
  if (dtype == 0) dgrad_t<float>(dy, lddy, w, dx, lddx, NB, H, W, C, stride); else dgrad_t<bf16>(dy, lddy, w, dx, lddx, NB, H, W, C, stride);
}
void emu_dw_wgrad(const void* x, long ldx, const void* dy, long lddy, float* dw, int NB, int H, int W, int C, int stride, int dtype, int gx) {
  if (dtype == 0) wgrad_t<float>(x, ldx, dy, lddy, dw, NB, H, W, C, stride, gx); else wgrad_t<bf16>(x, ldx, dy, lddy, dw, NB, H, W, C, stride, gx);
}
void emu_im2col3x3s2(const float* x, void* col, int NB, int H, int W, int Kp, int dtype) {
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long nt = (long)NB * Ho * Wo * (Kp >> 2);
  for (long t = 0; t < nt; ++t) { if (dtype == 0) dwc::im2col3x3s2<float, 3>(t, x, (float*)col, NB, H, W, Ho, Wo, Kp); else dwc::im2col3x3s2<bf16, 3>(t, x, (bf16*)col, NB, H, W, Ho, Wo, Kp); }
}
void emu_add_rows(const void* a, long lda, const void* b, long ldb, void* o, long ldo, long M, int C, int dtype) {
  const long nt = M * (C >> 2);
  for (long t = 0; t < nt; ++t) { if (dtype == 0) dwc::add_rows<float>(t, (const float*)a, lda, (const float*)b, ldb, (float*)o, ldo, M, C); else dwc::add_rows<bf16>(t, (const bf16*)a, lda, (const bf16*)b, ldb, (bf16*)o, ldo, M, C); }
}
}
