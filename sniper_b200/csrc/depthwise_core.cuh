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

// Border handling in all three kernels: every load is issued UNCONDITIONALLY from a clamped (always valid) address
// and the value is zeroed afterwards when the tap falls outside the map.  With the loads behind per-thread branches
// (first version: `if (outside) continue;`) the compiler could not hoist them and each warp had one load in flight at
// a time: 0.03-0.3 of the HBM roof (profiles/config4_r02.json).

// ---- forward: y[n,ho,wo,c] = sum_{kh,kw} x[n, ho*S+kh-1, wo*S+kw-1, c] * w[kh*3+kw, c]
// One thread = PW consecutive output pixels of one row x 4 channels: the (PW-1)*S+3 input columns of a row are loaded
// once (all of them in flight together) and feed every output they belong to (stride 1, PW 4: 18 loads for 4 outputs
// instead of 36).
template <typename T, int S, int PW>
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
  float acc[PW][4];
#pragma unroll
  for (int q = 0; q < PW; ++q) zero4(acc[q]);
  constexpr int NCOL = (PW - 1) * S + 3;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int hi = ho * S + kh - 1;
    const bool hv = hi >= 0 && hi < p.H;
    const T* row = x + ((n * p.H + clampi(hi, 0, p.H - 1)) * (long)p.W) * p.ldx + c;
    float v[NCOL][4], wk[3][4];
#pragma unroll
    for (int j = 0; j < NCOL; ++j) load4(row + (long)clampi(wo0 * S + j - 1, 0, p.W - 1) * p.ldx, v[j]);
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) load4(w + (long)(kh * 3 + kw) * p.C + c, wk[kw]);
#pragma unroll
    for (int j = 0; j < NCOL; ++j) {
      const int wi = wo0 * S + j - 1;
      if (!(hv && wi >= 0 && wi < p.W)) zero4(v[j]);
#pragma unroll
      for (int q = 0; q < PW; ++q) {
        const int kw = j - q * S;          // compile-time after unrolling
        if (kw >= 0 && kw < 3) {
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[q][k] = fmaf(v[j][k], wk[kw][k], acc[q][k]);
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

// ---- data gradient: dx[n,h,w,c] = sum_{kh,kw : (h+1-kh) % S == 0, (w+1-kw) % S == 0}
//                                   dy[n, (h+1-kh)/S, (w+1-kw)/S, c] * w[kh*3+kw, c]
// One thread = one input pixel x 4 channels (gather form: no atomics, deterministic).  Stride 1: 3 x 3 taps.  Stride 2:
// only the taps of matching parity exist -- kh = 1 for even h, kh in {0, 2} for odd h (same for w): at most 2 x 2 loads.
// Params describe the FORWARD convolution (H, W = dx map; Ho, Wo = dy map; ldx = dx pixel stride, ldy = dy pixel stride).
template <typename T, int S>
DW_HD void dgrad(long tid, const T* __restrict__ dy, const float* __restrict__ w, T* __restrict__ dx, const Params& p) {
  const int CV = p.C >> 2;
  long t = tid;
  const int cv = (int)(t % CV); t /= CV;
  const int wi = (int)(t % p.W); t /= p.W;
  const int hi = (int)(t % p.H);
  const long n = t / p.H;
  if (n >= p.NB) return;
  const int c = cv << 2;
  constexpr int NK = S == 1 ? 3 : 2;
  int khs[NK], kws[NK], hos[NK], wos[NK];
  bool hval[NK], wval[NK];
#pragma unroll
  for (int i = 0; i < NK; ++i) {
    khs[i] = S == 1 ? i : (i == 0 ? ((hi & 1) ? 0 : 1) : 2);
    kws[i] = S == 1 ? i : (i == 0 ? ((wi & 1) ? 0 : 1) : 2);
    const int a = hi + 1 - khs[i], b = wi + 1 - kws[i];          // = ho * S, wo * S when the tap exists
    hos[i] = a / S;
    wos[i] = b / S;
    hval[i] = a >= 0 && hos[i] < p.Ho && (S == 1 || i == 0 || (hi & 1));
    wval[i] = b >= 0 && wos[i] < p.Wo && (S == 1 || i == 0 || (wi & 1));
    hos[i] = clampi(hos[i], 0, p.Ho - 1);
    wos[i] = clampi(wos[i], 0, p.Wo - 1);
  }
  float g[NK][NK][4], wk[NK][NK][4];
#pragma unroll
  for (int i = 0; i < NK; ++i)
#pragma unroll
    for (int j = 0; j < NK; ++j) {
      load4(dy + ((n * p.Ho + hos[i]) * (long)p.Wo + wos[j]) * p.ldy + c, g[i][j]);
      load4(w + (long)(khs[i] * 3 + kws[j]) * p.C + c, wk[i][j]);
    }
  float acc[4];
  zero4(acc);
#pragma unroll
  for (int i = 0; i < NK; ++i)
#pragma unroll
    for (int j = 0; j < NK; ++j) {
      if (!(hval[i] && wval[j])) zero4(g[i][j]);
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = fmaf(g[i][j][k], wk[i][j][k], acc[k]);
    }
  store4(dx + ((n * p.H + hi) * (long)p.W + wi) * p.ldx + c, acc);
}
DW_HD long dgrad_threads(const Params& p) { return (long)p.NB * p.H * p.W * (p.C >> 2); }

// ---- weight gradient, per-thread partial: acc[t][k] = sum over this thread's output pixels of
//      dy[q, c+k] * x[n, ho*S+kh-1, wo*S+kw-1, c+k].
// Block = 32 x TY threads.  A warp row (32 lanes) covers LC channel vectors x 32/LC pixels, LC = 32 when C/4 is a
// multiple of 32, else 16 (C is a multiple of 64): no idle lanes for C = 64 / 192 / 320 / 576 / 960.  Thread (tx, ty) of
// block (bx, by): channel vector by*LC + tx%LC, pixel lane ty*(32/LC) + tx/LC of TY*(32/LC); output pixels
// q = bx*PL + lane, + gridx*PL, ...  (neighbouring lanes take neighbouring pixels: their windows overlap in L1).
// Returns the first channel of the thread, or -1 when it is beyond C.
DW_HD int wgrad_lc(int C) { return ((C >> 2) % 32 == 0) ? 32 : 16; }
template <typename T, int S>
DW_HD int wgrad_partial(int bx, int by, int tx, int ty, int TY, int gridx, const T* __restrict__ x,
                        const T* __restrict__ dy, const Params& p, float (&acc)[9][4]) {
#pragma unroll
  for (int t = 0; t < 9; ++t) zero4(acc[t]);
  const int LC = wgrad_lc(p.C), PPW = 32 / LC;
  const int c = (by * LC + (tx % LC)) << 2;
  if (c >= p.C) return -1;
  const int PL = TY * PPW;
  const long total = (long)p.NB * p.Ho * p.Wo;
  for (long q = (long)bx * PL + ty * PPW + tx / LC; q < total; q += (long)gridx * PL) {
    const int wo = (int)(q % p.Wo);
    const long r = q / p.Wo;
    const int ho = (int)(r % p.Ho);
    const long n = r / p.Ho;
    float g[4], v[9][4];
    load4(dy + q * p.ldy + c, g);
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const T* row = x + ((n * p.H + clampi(ho * S + kh - 1, 0, p.H - 1)) * (long)p.W) * p.ldx + c;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) load4(row + (long)clampi(wo * S + kw - 1, 0, p.W - 1) * p.ldx, v[kh * 3 + kw]);
    }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int hi = ho * S + kh - 1;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int wi = wo * S + kw - 1;
        if (!(hi >= 0 && hi < p.H && wi >= 0 && wi < p.W)) zero4(v[kh * 3 + kw]);
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[kh * 3 + kw][k] = fmaf(g[k], v[kh * 3 + kw][k], acc[kh * 3 + kw][k]);
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
