// Batched per-(image, class) soft-NMS / NMS on the device for the multi-scale inference aggregation.
//
// Replaces the `Pool(32)` of host processes running `cpu_soft_nms` (lib/nms/cpu_nms.pyx:17-110, Gaussian, sigma 0.55,
// score threshold 0.001; called per image and class by Tester.aggregate, lib/inference.py:152-200, through
// nms_worker.worker) -- the "per-scale NMS latency" of BASELINE config 5.  One CTA per problem; all problems of a scale
// (images x 80 classes) run in one launch.
//
// Semantics of the reference loop, restated order-free: repeat { take the highest-scoring remaining box (rows the
// reference has swapped to the front are exactly the boxes taken so far, in order); multiply the score of every other
// remaining box by w(IoU with it); drop boxes whose score falls below `threshold` }.  The output rows are the taken boxes
// in the order taken, each with the score it had when taken = the reference's returned array.  Arithmetic follows the
// Cythonized reference (float IoU with the `+ 1` sums evaluated in double and narrowed once, Gaussian weight =
// exp evaluated in double); this file is compiled with -fmad=false.  Exact score ties are broken by the lower input row
// (the reference breaks them by its current, swap-dependent order).
#include "common.cuh"
#include <math.h>

namespace {

constexpr int kNmsTPB = 256;

__global__ void __launch_bounds__(kNmsTPB) soft_nms_batched_kernel(float* __restrict__ dets, const int* __restrict__ offsets,
                                                                   float sigma, float Nt, float threshold,
                                                                   unsigned method, int* __restrict__ out_counts,
                                                                   float* __restrict__ scratch) {
  const int p = blockIdx.x;
  const int base = offsets[p], N = offsets[p + 1] - base;
  float* d = dets + (size_t)base * 5;
  float* work = scratch + (size_t)base * 5;     // the remaining boxes (score < 0 marks a removed / taken row)
  __shared__ float s_val[kNmsTPB / 32];
  __shared__ int s_idx[kNmsTPB / 32];
  __shared__ float s_t[5];
  __shared__ int s_stop;
  for (int i = threadIdx.x; i < N * 5; i += blockDim.x) work[i] = d[i];
  __syncthreads();
  int taken = 0;
  for (;;) {
    // ---- argmax over the remaining rows (scores are >= threshold > 0 or original; removed rows carry -1)
    float best = -1.f;
    int bi = -1;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      const float s = work[5 * j + 4];
      if (s > best) { best = s; bi = j; }          // strict >: the lowest row index wins a tie within a thread
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ob > best || (ob == best && oi >= 0 && (bi < 0 || oi < bi))) { best = ob; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float b = s_val[0];
      int k = s_idx[0];
      for (int w = 1; w < kNmsTPB / 32; ++w)
        if (s_val[w] > b || (s_val[w] == b && s_idx[w] >= 0 && (k < 0 || s_idx[w] < k))) { b = s_val[w]; k = s_idx[w]; }
      s_stop = (k < 0 || b < 0.f) ? 1 : 0;
      if (!s_stop) {
#pragma unroll
        for (int c = 0; c < 5; ++c) s_t[c] = work[5 * k + c];
        work[5 * k + 4] = -1.f;                    // taken
#pragma unroll
        for (int c = 0; c < 5; ++c) d[5 * taken + c] = s_t[c];
      }
    }
    __syncthreads();
    if (s_stop) break;
    ++taken;
    const float tx1 = s_t[0], ty1 = s_t[1], tx2 = s_t[2], ty2 = s_t[3];
    // ---- decay / remove the others
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      const float s = work[5 * j + 4];
      if (s < 0.f) continue;
      const float x1 = work[5 * j], y1 = work[5 * j + 1], x2 = work[5 * j + 2], y2 = work[5 * j + 3];
      const float area = (float)(((double)(x2 - x1) + 1.0) * ((double)(y2 - y1) + 1.0));
      const float iw = (float)((double)(fminf(tx2, x2) - fmaxf(tx1, x1)) + 1.0);
      if (iw > 0) {
        const float ih = (float)((double)(fminf(ty2, y2) - fmaxf(ty1, y1)) + 1.0);
        if (ih > 0) {
          const float inter = iw * ih;
          const float ua = (float)(((((double)(tx2 - tx1) + 1.0) * ((double)(ty2 - ty1) + 1.0)) + (double)area) -
                                   (double)inter);
          const float ov = inter / ua;
          float weight;
          if (method == 1) weight = ov > Nt ? (float)(1.0 - (double)ov) : 1.f;
          else if (method == 2) weight = (float)exp((double)(-(ov * ov) / sigma));
          else weight = ov > Nt ? 0.f : 1.f;
          const float ns = weight * s;
          work[5 * j + 4] = ns < threshold ? -1.f : ns;
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) out_counts[p] = taken;
}

}  // namespace

extern "C" {

// dets: [total, 5] (x1, y1, x2, y2, score) rows of all problems back to back, problem p = rows [offsets[p],
// offsets[p+1]); rewritten in place: the first out_counts[p] rows of each segment are the surviving detections in the
// order cpu_soft_nms returns them.  offsets: device int32 [P+1]; scratch: device float [total*5].  Scores must be >= 0.
// method 1 = linear, 2 = Gaussian (the reference's TEST default), 3 = hard NMS.
int sniper_soft_nms_batched(float* dets, const int* offsets, int P, float sigma, float Nt, float threshold,
                            unsigned method, int* out_counts, float* scratch, void* stream) {
  if (P <= 0) return 0;
  SN_CHECK(method >= 1 && method <= 3, "soft_nms_batched: method must be 1 (linear), 2 (gaussian) or 3 (hard)");
  SN_CHECK(threshold >= 0.f, "soft_nms_batched: threshold must be >= 0");
  soft_nms_batched_kernel<<<P, kNmsTPB, 0, (cudaStream_t)stream>>>(dets, offsets, sigma, Nt, threshold, method,
                                                                   out_counts, scratch);
  SN_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
