// Per-thread bodies of the depthwise 3x3 kernels (depthwise.cu), written as __host__ __device__ functions of the
// linear thread index so that tests/dw_emulate.cu can run exactly this index arithmetic on the CPU (this container has
// no GPU: the emulation pins the addressing / tap algebra against torch here, the -m gpu tests pin the launches).
//
// Operator: mx.sym.Convolution(kernel=(3,3), pad=(1,1), stride=(s,s), num_group=C, num_filter=C, no_bias=True) of
// mobilenetv2_e2e.py:58-66 (reference kernels: src/operator/nn/depthwise_convolution-inl.h / depthwise_convolution_tf.cuh).
// Layout: NHWC activations ([pixels, C] rows with a row stride), weights tap-major [9, C] fp32 (the reference's
// (C,1,3,3) transposed), 4 channels per thread (16 bytes of fp32, 8 bytes of bf16), fp32 accumulation.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>
#include <string.h>

#define DW_HD __host__ __device__ __forceinline__
// (scheduling of the loads: see `join_loads` below)

namespace dwc {

struct Params {
  int NB, H, W, C;       // input map
  int Ho, Wo, stride;    // output map: Ho = (H + 2 - 3) / stride + 1
  long ldx, ldy;         // elements between consecutive pixels of the input / output map
};

DW_HD float bits2f(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
DW_HD uint32_t f2bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
// round-to-nearest-even fp32 -> bf16 bits (== __float2bfloat16_rn for finite inputs; NaN stays NaN)
DW_HD uint32_t f2bf(float f) {
  uint32_t x = f2bits(f);
  if ((x & 0x7fffffffu) > 0x7f800000u) return (x >> 16) | 0x40u;
  x += 0x7fffu + ((x >> 16) & 1u);
  return x >> 16;
}

DW_HD void load4(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
DW_HD void load4(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  v[0] = bits2f(u.x << 16); v[1] = bits2f(u.x & 0xffff0000u);
  v[2] = bits2f(u.y << 16); v[3] = bits2f(u.y & 0xffff0000u);
}
DW_HD void store4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
DW_HD void store4(__nv_bfloat16* p, const float (&v)[4]) {
  uint2 u;
  u.x = f2bf(v[0]) | (f2bf(v[1]) << 16);
  u.y = f2bf(v[2]) | (f2bf(v[3]) << 16);
  *reinterpret_cast<uint2*>(p) = u;
}

DW_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
DW_HD void zero4(float (&v)[4]) { v[0] = v[1] = v[2] = v[3] = 0.f; }

// Packed form of a 4-channel vector as it sits in memory (16 bytes fp32, 8 bytes bf16): loads are kept packed until they
// are consumed, so that a thread can hold a whole 3 x 6 window in flight (18 loads = 36 registers in bf16).
template <typename T> struct Raw;
template <> struct Raw<float> { typedef float4 type; };
template <> struct Raw<__nv_bfloat16> { typedef uint2 type; };
DW_HD float4 loadraw(const float* p) { return *reinterpret_cast<const float4*>(p); }
DW_HD uint2 loadraw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint2*>(p); }
DW_HD void unpack(const float4& t, float (&v)[4]) { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
DW_HD void unpack(const uint2& u, float (&v)[4]) {
  v[0] = bits2f(u.x << 16); v[1] = bits2f(u.x & 0xffff0000u);
  v[2] = bits2f(u.y << 16); v[3] = bits2f(u.y & 0xffff0000u);
}

// join_loads: how "issue EVERY load of the window, then consume" is enforced.  Left alone, ptxas sinks loads towards
// their first use to save registers, so a warp pays the memory latency once per small group of loads (5-8 dependent
// stalls per thread in these kernels; a warp-level fence is optimised away, a CTA-level one costs a MEMBAR).  Here
// every packed load is folded into one checksum word and the accumulators are INITIALISED with a value that depends on
// it: `(m == magic) & (p.NB < 0) ? 1 : 0` is always 0 at run time (the launcher rejects NB <= 0) but not provably so at
// compile time.  Every FMA chain therefore starts after all loads of the window have returned -- one latency per
// thread instead of one per group -- at the price of ~20 integer instructions.  Results are unchanged (the host
// emulation executes the same expression).
DW_HD uint32_t fold(uint32_t m, const float4& r) { return m ^ f2bits(r.x) ^ f2bits(r.y) ^ f2bits(r.z) ^ f2bits(r.w); }
DW_HD uint32_t fold(uint32_t m, const uint2& r) { return m ^ r.x ^ r.y; }
DW_HD float join_loads(uint32_t m, const Params& p) { return ((m == 0x9e3779b9u) & (p.NB < 0)) ? 1.f : 0.f; }

// Border handling and load scheduling in all kernels: every load is issued UNCONDITIONALLY from a clamped (always valid)
// address, ALL loads of a thread's window are issued before the first use, and values of taps outside the map are zeroed
// afterwards.  History (profiles/config4_r02.md): loads behind per-thread branches (`if (outside) continue;`) could not be
// hoisted -> one load in flight per warp, 0.03-0.3 of the HBM roof; unconditional loads issued row by row -> the forward
// kernel sat exactly at (resident warps x one row of loads) / latency = 1.5 TB/s.

// ---- forward: y[n,ho,wo,c] = sum_{kh,kw} x[n, ho*S+kh-1, wo*S+kw-1, c] * w[kh*3+kw, c]
// One thread = PW consecutive output pixels of one row x 4 channels: the 3 x ((PW-1)*S+3) input window is loaded once
// and feeds every output it belongs to (stride 1, PW 4: 18 loads for 4 outputs instead of 36).  FLIP: taps taken in
// reverse order (w[8 - t]) -- the stride-1 DATA GRADIENT is this kernel run on dy with the flipped filter.
template <typename T, int S, int PW, bool FLIP>
DW_HD void fwd(long tid, const T* __restrict__ x, const float* __restrict__ w, T* __restrict__ y, const Params& p) {
  const int CV = p.C >> 2;
  const int WB = (p.Wo + PW - 1) / PW;
  long t = tid;
  const int cv = (int)(t % CV); t /= CV;
  const int wb = (int)(t % WB); t /= WB;
  const int ho = (int)(t % p.Ho);
  const long n = t / p.Ho;
  if (n >= p.NB) return;
  const int c = cv << 2;
  const int wo0 = wb * PW;
  constexpr int NCOL = (PW - 1) * S + 3;
  typename Raw<T>::type raw[3][NCOL];
  float wk[9][4];
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const T* row = x + ((n * p.H + clampi(ho * S + kh - 1, 0, p.H - 1)) * (long)p.W) * p.ldx + c;
#pragma unroll
    for (int j = 0; j < NCOL; ++j) raw[kh][j] = loadraw(row + (long)clampi(wo0 * S + j - 1, 0, p.W - 1) * p.ldx);
  }
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) load4(w + (long)(FLIP ? 8 - tp : tp) * p.C + c, wk[tp]);
  uint32_t m = 0;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int j = 0; j < NCOL; ++j) m = fold(m, raw[kh][j]);
  const float a0 = join_loads(m, p);
  float acc[PW][4];
#pragma unroll
  for (int q = 0; q < PW; ++q) { acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = a0; }
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int hi = ho * S + kh - 1;
    const bool hv = hi >= 0 && hi < p.H;
#pragma unroll
    for (int j = 0; j < NCOL; ++j) {
      const int wi = wo0 * S + j - 1;
      float v[4];
      unpack(raw[kh][j], v);
      if (!(hv && wi >= 0 && wi < p.W)) zero4(v);
#pragma unroll
      for (int q = 0; q < PW; ++q) {
        const int kw = j - q * S;          // compile-time after unrolling
        if (kw >= 0 && kw < 3) {
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[q][k] = fmaf(v[k], wk[kh * 3 + kw][k], acc[q][k]);
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < PW; ++q)
    if (wo0 + q < p.Wo) store4(y + ((n * p.Ho + ho) * (long)p.Wo + wo0 + q) * p.ldy + c, acc[q]);
}
template <int S, int PW>
DW_HD long fwd_threads(const Params& p) {
  return (long)p.NB * p.Ho * ((p.Wo + PW - 1) / PW) * (p.C >> 2);
}

// ---- forward, shared-memory tiled (bf16, C a multiple of 64): a block stages the (TH-1)*S+3 x (TW-1)*S+3 input window
// of a TH x TW output tile for one 64-channel chunk (128 bytes per pixel) with 16-byte asynchronous copies (cp.async,
// zero fill outside the map: every copy of the block is in flight at once, each input byte is fetched 1.16-1.33 times
// instead of 4.5 times through L1), then every thread computes strips of PW outputs from shared memory.  FLIP as above.
// Split into the two phases so that the host emulation can run them block by block.
template <int S> struct Tile {
  static constexpr int TH = S == 1 ? 8 : 4, TW = S == 1 ? 32 : 16, PW = S == 1 ? 4 : 2;
  static constexpr int IH = (TH - 1) * S + 3, IW = (TW - 1) * S + 3;
  static constexpr int ELEMS = IH * IW * 64;               // bf16 elements of the staged window (43.5 KB / 38 KB)
};
struct TileId { long n; int ho0, wo0, c0; };
template <int S>
DW_HD TileId tile_id(long bx, int by, const Params& p) {
  const int tx = (p.Wo + Tile<S>::TW - 1) / Tile<S>::TW, ty = (p.Ho + Tile<S>::TH - 1) / Tile<S>::TH;
  TileId t;
  t.wo0 = (int)(bx % tx) * Tile<S>::TW; bx /= tx;
  t.ho0 = (int)(bx % ty) * Tile<S>::TH;
  t.n = bx / ty;
  t.c0 = by * 64;
  return t;
}
template <int S>
DW_HD long tile_blocks(const Params& p) {
  return (long)p.NB * ((p.Wo + Tile<S>::TW - 1) / Tile<S>::TW) * ((p.Ho + Tile<S>::TH - 1) / Tile<S>::TH);
}
// 16 bytes global -> shared, or 16 zero bytes when !valid (src must be a valid address either way)
DW_HD void copy16(__nv_bfloat16* dst, const __nv_bfloat16* src, bool valid) {
#ifdef __CUDA_ARCH__
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
#else
  if (valid) memcpy(dst, src, 16); else memset(dst, 0, 16);
#endif
}
template <int S>
DW_HD void tile_stage(int tid, int nthreads, const TileId& t, const __nv_bfloat16* __restrict__ x, __nv_bfloat16* smem,
                      const Params& p) {
  constexpr int IW = Tile<S>::IW, CH = Tile<S>::IH * IW * 8;        // 16-byte pieces: 8 per pixel
  for (int i = tid; i < CH; i += nthreads) {
    const int pix = i >> 3, part = i & 7;
    const int ih = pix / IW, iw = pix - ih * IW;
    const int hi = t.ho0 * S - 1 + ih, wi = t.wo0 * S - 1 + iw;
    const bool valid = hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
    const __nv_bfloat16* src = x + ((t.n * p.H + clampi(hi, 0, p.H - 1)) * (long)p.W + clampi(wi, 0, p.W - 1)) * p.ldx +
                               t.c0 + part * 8;
    copy16(smem + pix * 64 + part * 8, src, valid);
  }
}
template <int S, bool FLIP>
DW_HD void tile_compute(int tid, int nthreads, const TileId& t, const __nv_bfloat16* smem, const float* __restrict__ w,
                        __nv_bfloat16* __restrict__ y, const Params& p) {
  constexpr int TW = Tile<S>::TW, TH = Tile<S>::TH, PW = Tile<S>::PW, IW = Tile<S>::IW, SB = TW / PW;
  constexpr int NCOL = (PW - 1) * S + 3;
  const int cv = tid & 15;                                  // nthreads is a multiple of 16: fixed per thread
  const int c = t.c0 + (cv << 2);
  float wk[9][4];
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) load4(w + (long)(FLIP ? 8 - tp : tp) * p.C + c, wk[tp]);
  for (int s = tid; s < TH * SB * 16; s += nthreads) {
    const int sb = (s >> 4) % SB, r = (s >> 4) / SB;
    const int ho = t.ho0 + r, wo0 = t.wo0 + sb * PW;
    if (ho >= p.Ho || wo0 >= p.Wo) continue;
    float acc[PW][4];
#pragma unroll
    for (int q = 0; q < PW; ++q) zero4(acc[q]);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const __nv_bfloat16* row = smem + ((r * S + kh) * IW + sb * PW * S) * 64 + (cv << 2);
#pragma unroll
      for (int j = 0; j < NCOL; ++j) {
        float v[4];
        load4(row + j * 64, v);
#pragma unroll
        for (int q = 0; q < PW; ++q) {
          const int kw = j - q * S;
          if (kw >= 0 && kw < 3) {
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[q][k] = fmaf(v[k], wk[kh * 3 + kw][k], acc[q][k]);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < PW; ++q)
      if (wo0 + q < p.Wo) store4(y + ((t.n * p.Ho + ho) * (long)p.Wo + wo0 + q) * p.ldy + c, acc[q]);
  }
}

// ---- data gradient, stride 2: one thread = one dy pixel (a, b) x 4 channels -> the 2 x 2 block of dx at (2a + i, 2b + j).
// With x index h = 2*ho + kh - 1 only taps of matching parity exist:
//   dx[2a  , 2b  ] = dy[a][b] w11
//   dx[2a  , 2b+1] = dy[a][b+1] w10 + dy[a][b] w12
//   dx[2a+1, 2b  ] = dy[a+1][b] w01 + dy[a][b] w21
//   dx[2a+1, 2b+1] = dy[a+1][b+1] w00 + dy[a+1][b] w02 + dy[a][b+1] w20 + dy[a][b] w22
// 4 dy loads + 9 weight loads for 4 outputs (the gather-per-input-pixel form needed up to 4 + 4 loads per output).
// Params describe the FORWARD convolution (H, W = dx map; Ho, Wo = dy map; ldx = dx pixel stride, ldy = dy pixel stride).
// Stride 1 uses fwd<T, 1, 4, true> on dy.
template <typename T>
DW_HD void dgrad_s2(long tid, const T* __restrict__ dy, const float* __restrict__ w, T* __restrict__ dx, const Params& p) {
  const int CV = p.C >> 2;
  long t = tid;
  const int cv = (int)(t % CV); t /= CV;
  const int b = (int)(t % p.Wo); t /= p.Wo;
  const int a = (int)(t % p.Ho);
  const long n = t / p.Ho;
  if (n >= p.NB) return;
  const int c = cv << 2;
  const bool a1 = a + 1 < p.Ho, b1 = b + 1 < p.Wo;
  const T* r0 = dy + ((n * p.Ho + a) * (long)p.Wo) * p.ldy + c;
  const T* r1 = dy + ((n * p.Ho + (a1 ? a + 1 : a)) * (long)p.Wo) * p.ldy + c;
  const long o0 = (long)b * p.ldy, o1 = (long)(b1 ? b + 1 : b) * p.ldy;
  typename Raw<T>::type q00 = loadraw(r0 + o0), q01 = loadraw(r0 + o1), q10 = loadraw(r1 + o0), q11 = loadraw(r1 + o1);
  float wk[9][4];
#pragma unroll
  for (int tp = 0; tp < 9; ++tp) load4(w + (long)tp * p.C + c, wk[tp]);
  float g00[4], g01[4], g10[4], g11[4];
  unpack(q00, g00); unpack(q01, g01); unpack(q10, g10); unpack(q11, g11);
  if (!b1) zero4(g01);
  if (!a1) zero4(g10);
  if (!(a1 && b1)) zero4(g11);
  float o[2][2][4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    o[0][0][k] = g00[k] * wk[4][k];
    o[0][1][k] = fmaf(g01[k], wk[3][k], g00[k] * wk[5][k]);
    o[1][0][k] = fmaf(g10[k], wk[1][k], g00[k] * wk[7][k]);
    o[1][1][k] = fmaf(g11[k], wk[0][k], fmaf(g10[k], wk[2][k], fmaf(g01[k], wk[6][k], g00[k] * wk[8][k])));
  }
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int h = 2 * a + i, ww = 2 * b + j;
      if (h < p.H && ww < p.W) store4(dx + ((n * p.H + h) * (long)p.W + ww) * p.ldx + c, o[i][j]);
    }
}
DW_HD long dgrad_s2_threads(const Params& p) { return (long)p.NB * p.Ho * p.Wo * (p.C >> 2); }

// ---- weight gradient, per-thread partial: acc[t][k] = sum over this thread's output pixels of
//      dy[q, c+k] * x[n, ho*S+kh-1, wo*S+kw-1, c+k].
// Block = 32 x TY threads.  A warp row (32 lanes) covers LC channel vectors x 32/LC strips, LC = 32 when C/4 is a
// multiple of 32, else 16 (C is a multiple of 64): no idle lanes for C = 64 / 192 / 320 / 576 / 960.  A thread walks over
// STRIPS of PW consecutive output pixels of a row (3 x ((PW-1)*S+3) input window + PW dy vectors, all in flight
// together); thread (tx, ty) of block (bx, by): channel vector by*LC + tx%LC, strip lane ty*(32/LC) + tx/LC of
// PL = TY*(32/LC); strips bx*PL + lane, + gridx*PL, ...  The result does not depend on gridx (any grid gives the same
// sums up to fp32 association).  Returns the first channel of the thread, or -1 when it is beyond C.
DW_HD int wgrad_lc(int C) { return ((C >> 2) % 32 == 0) ? 32 : 16; }
template <typename T, int S, int PW>
DW_HD int wgrad_partial(int bx, int by, int tx, int ty, int TY, int gridx, const T* __restrict__ x,
                        const T* __restrict__ dy, const Params& p, float (&acc)[9][4]) {
#pragma unroll
  for (int t = 0; t < 9; ++t) zero4(acc[t]);
  const int LC = wgrad_lc(p.C), PPW = 32 / LC;
  const int c = (by * LC + (tx % LC)) << 2;
  if (c >= p.C) return -1;
  const int PL = TY * PPW;
  const int WB = (p.Wo + PW - 1) / PW;
  const long total = (long)p.NB * p.Ho * WB;
  constexpr int NCOL = (PW - 1) * S + 3;
  for (long sidx = (long)bx * PL + ty * PPW + tx / LC; sidx < total; sidx += (long)gridx * PL) {
    const int wo0 = (int)(sidx % WB) * PW;
    const long r = sidx / WB;
    const int ho = (int)(r % p.Ho);
    const long n = r / p.Ho;
    typename Raw<T>::type rg[PW], rx[3][NCOL];
    const T* grow = dy + ((n * p.Ho + ho) * (long)p.Wo) * p.ldy + c;
#pragma unroll
    for (int q = 0; q < PW; ++q) rg[q] = loadraw(grow + (long)clampi(wo0 + q, 0, p.Wo - 1) * p.ldy);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const T* row = x + ((n * p.H + clampi(ho * S + kh - 1, 0, p.H - 1)) * (long)p.W) * p.ldx + c;
#pragma unroll
      for (int j = 0; j < NCOL; ++j) rx[kh][j] = loadraw(row + (long)clampi(wo0 * S + j - 1, 0, p.W - 1) * p.ldx);
    }
      uint32_t m = 0;
#pragma unroll
    for (int q = 0; q < PW; ++q) m = fold(m, rg[q]);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int j = 0; j < NCOL; ++j) m = fold(m, rx[kh][j]);
    const float a0 = join_loads(m, p);
    float g[PW][4];
#pragma unroll
    for (int q = 0; q < PW; ++q) {
      unpack(rg[q], g[q]);
      if (wo0 + q >= p.Wo) zero4(g[q]);
#pragma unroll
      for (int k = 0; k < 4; ++k) g[q][k] += a0;          // every product of this strip waits for the whole window
    }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int hi = ho * S + kh - 1;
      const bool hv = hi >= 0 && hi < p.H;
#pragma unroll
      for (int j = 0; j < NCOL; ++j) {
        const int wi = wo0 * S + j - 1;
        float v[4];
        unpack(rx[kh][j], v);
        if (!(hv && wi >= 0 && wi < p.W)) zero4(v);
#pragma unroll
        for (int q = 0; q < PW; ++q) {
          const int kw = j - q * S;
          if (kw >= 0 && kw < 3) {
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[kh * 3 + kw][k] = fmaf(g[q][k], v[k], acc[kh * 3 + kw][k]);
          }
        }
      }
    }
  }
  return c;
}

// ---- im2col of the first layer: 3x3 / stride 2 / pad 1 over an fp32 NCHW image with CIN channels
// (mobilenetv2_e2e.py:204-212 'first-3x3-conv').  col[(n,ho,wo), k] with k = (kh*3+kw)*CIN + ci for k < 9*CIN, zero
// up to Kp; one thread = 4 consecutive k of one output pixel.
template <typename T, int CIN>
DW_HD void im2col3x3s2(long tid, const float* __restrict__ x, T* __restrict__ col, int NB, int H, int W, int Ho, int Wo,
                       int Kp) {
  const int KV = Kp >> 2;
  long t = tid;
  const int kv = (int)(t % KV); t /= KV;
  const int wo = (int)(t % Wo); t /= Wo;
  const int ho = (int)(t % Ho);
  const long n = t / Ho;
  if (n >= NB) return;
  float v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int k = kv * 4 + e;
    float val = 0.f;
    if (k < 9 * CIN) {
      const int tap = k / CIN, ci = k - tap * CIN;
      const int hi = ho * 2 + tap / 3 - 1, wi = wo * 2 + tap % 3 - 1;
      if (hi >= 0 && hi < H && wi >= 0 && wi < W) val = x[((n * CIN + ci) * H + hi) * (long)W + wi];
    }
    v[e] = val;
  }
  store4(col + ((n * Ho + ho) * (long)Wo + wo) * Kp + kv * 4, v);
}

// ---- out = a + b on [M, C] rows (elemwise_add of the inverted-residual shortcut, mobilenetv2_e2e.py:22-24)
template <typename T>
DW_HD void add_rows(long tid, const T* a, long lda, const T* b, long ldb, T* o,
                    long ldo, long M, int C) {
  const int CV = C >> 2;
  const long r = tid / CV;
  const int c = (int)(tid % CV) << 2;
  if (r >= M) return;
  float va[4], vb[4];
  load4(a + r * lda + c, va);
  load4(b + r * ldb + c, vb);
#pragma unroll
  for (int k = 0; k < 4; ++k) va[k] += vb[k];
  store4(o + r * ldo + c, va);
}

}  // namespace dwc
