// RPN anchor matching on the device (north_star: "anchor matching ... coalesced/vectorised HBM kernels").
//
// Replaces the per-chip host work of anchor_worker.worker (lib/data_utils/data_workers.py:194-363):
// anchor grid (:141-155, generate_anchor.py:8-77), inside test (:198-201), float64 IoU with the valid and
// the invalid GT sets (lib/bbox/bbox.pyx:17-57), label rules (:300-318), bbox_transform targets
// (lib/bbox/bbox_transform.py:64-90) and the (a,h,w)/(4a,h,w) packing (:346-356).  The random fg/bg
// subsampling (:326-338, npr.choice) stays a host decision: it arrives as an optional per-anchor disable mask.
// Arithmetic is float64 in the reference's operation order (this TU is built with -fmad=false).
#include "common.cuh"
#include <math.h>

namespace {

struct AtArgs {
  const float* gt;       // [B,G,4] valid GT boxes (first ngt[b] rows used)
  const int* ngt;        // [B]
  const float* inv;      // [B,Gi,4] invalid GT boxes
  const int* ninv;       // [B]
  const float* im_info;  // [B,3]
  const uint8_t* disable;  // optional [B,H*W*A] in flat (h,w,a) order
  int B, G, Gi, H, W, A, stride;
  double pos_thresh, neg_thresh;
  double base[4 * 64];
  unsigned long long* gtmax;  // [B,G] bit patterns of non-negative doubles (zero-init)
  float* label;          // [B,A*H*W]   (a,h,w)
  float* target;         // [B,4A,H,W]
  float* weight;         // [B,4A,H,W]
  int32_t* argmax;       // optional [B,H*W*A] flat (h,w,a), -1 outside
};

__device__ __forceinline__ void anchor_box(const AtArgs& p, int hw, int a, double* bx) {
  const int h = hw / p.W, w = hw - h * p.W;
  bx[0] = p.base[4 * a] + (double)(w * p.stride);
  bx[1] = p.base[4 * a + 1] + (double)(h * p.stride);
  bx[2] = p.base[4 * a + 2] + (double)(w * p.stride);
  bx[3] = p.base[4 * a + 3] + (double)(h * p.stride);
}

__device__ __forceinline__ bool inside(const AtArgs& p, int b, const double* bx) {
  // data_workers.py:198-201 (x2 is tested against im_info[0], y2 against im_info[1], as in the reference)
  return bx[0] >= -32 && bx[1] >= -32 && bx[2] < (double)p.im_info[3 * b] + 32 && bx[3] < (double)p.im_info[3 * b + 1] + 32;
}

__device__ __forceinline__ double iou(const double* bx, const float* q) {
  const double q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
  const double qa = (q2 - q0 + 1) * (q3 - q1 + 1);
  const double iw = fmin(bx[2], q2) - fmax(bx[0], q0) + 1;
  if (!(iw > 0)) return 0.0;
  const double ih = fmin(bx[3], q3) - fmax(bx[1], q1) + 1;
  if (!(ih > 0)) return 0.0;
  const double ua = (bx[2] - bx[0] + 1) * (bx[3] - bx[1] + 1) + qa - iw * ih;
  return iw * ih / ua;
}

__global__ void __launch_bounds__(256) at_gtmax_kernel(AtArgs p) {
  const int total = p.H * p.W * p.A;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= total) return;
  const int hw = idx / p.A, a = idx - hw * p.A;
  double bx[4];
  anchor_box(p, hw, a, bx);
  if (!inside(p, b, bx)) return;
  const int n = p.ngt[b];
  for (int g = 0; g < n; ++g) {
    const double o = iou(bx, p.gt + ((size_t)b * p.G + g) * 4);
    if (o > 0) atomicMax(p.gtmax + (size_t)b * p.G + g, (unsigned long long)__double_as_longlong(o));
  }
}

__global__ void __launch_bounds__(256) at_label_kernel(AtArgs p) {
  const int HW = p.H * p.W;
  const int total = HW * p.A;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (idx >= total) return;
  const int hw = idx / p.A, a = idx - hw * p.A;
  double bx[4];
  anchor_box(p, hw, a, bx);
  float lab = -1.f;
  float tg[4] = {0.f, 0.f, 0.f, 0.f};
  int am = -1;
  if (inside(p, b, bx)) {
    const int n = p.ngt[b], ni = p.ninv[b];
    double maxn = 0.0;
    for (int g = 0; g < ni; ++g) maxn = fmax(maxn, iou(bx, p.inv + ((size_t)b * p.Gi + g) * 4));
    if (n > 0) {
      double mx = -1.0;
      bool is_gt_argmax = false;
      for (int g = 0; g < n; ++g) {
        const double o = iou(bx, p.gt + ((size_t)b * p.G + g) * 4);
        if (o > mx) { mx = o; am = g; }   // np.argmax: first maximum
        const double gm = __longlong_as_double((long long)p.gtmax[(size_t)b * p.G + g]);
        is_gt_argmax |= (o == gm);        // np.where(overlaps == gt_max_overlaps): every tie, incl. gt_max == 0
      }
      if (mx < p.neg_thresh) lab = 0.f;
      if (is_gt_argmax) lab = 1.f;
      if (mx >= p.pos_thresh) lab = 1.f;
      if (ni > 0 && maxn > 0.3) lab = -1.f;
      const float* q = p.gt + ((size_t)b * p.G + am) * 4;
      const double ew = bx[2] - bx[0] + 1.0, eh = bx[3] - bx[1] + 1.0;
      const double ecx = bx[0] + 0.5 * (ew - 1.0), ecy = bx[1] + 0.5 * (eh - 1.0);
      const double gw = (double)q[2] - (double)q[0] + 1.0, gh = (double)q[3] - (double)q[1] + 1.0;
      const double gcx = (double)q[0] + 0.5 * (gw - 1.0), gcy = (double)q[1] + 0.5 * (gh - 1.0);
      tg[0] = (float)((gcx - ecx) / (ew + 1e-7));
      tg[1] = (float)((gcy - ecy) / (eh + 1e-7));
      tg[2] = (float)log(gw / (ew + 1e-7));
      tg[3] = (float)log(gh / (eh + 1e-7));
    } else {
      lab = 0.f;
      if (ni > 0 && maxn > 0.3) lab = -1.f;
    }
  }
  if (p.argmax) p.argmax[(size_t)b * total + idx] = am;
  if (p.disable && p.disable[(size_t)b * total + idx]) lab = -1.f;
  p.label[(size_t)b * total + (size_t)a * HW + hw] = lab;
  const float w = lab == 1.f ? 1.f : 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const size_t o = ((size_t)b * 4 * p.A + 4 * a + j) * HW + hw;
    p.target[o] = w == 1.f ? tg[j] : 0.f;
    p.weight[o] = w;
  }
}

// generate_anchor.py:8-77 in float64 (np.round == rint, half to even)
void make_base_anchors(int base_size, const float* ratios, int nr, const float* scales, int ns, double* out) {
  const double w0 = base_size, h0 = base_size, xc0 = 0.5 * (w0 - 1), yc0 = 0.5 * (h0 - 1);
  int n = 0;
  for (int i = 0; i < nr; ++i) {
    const double size_ratio = (w0 * h0) / (double)ratios[i];
    const double ws = rint(sqrt(size_ratio));
    const double hs = rint(ws * (double)ratios[i]);
    const double r0 = xc0 - 0.5 * (ws - 1), r1 = yc0 - 0.5 * (hs - 1), r2 = xc0 + 0.5 * (ws - 1), r3 = yc0 + 0.5 * (hs - 1);
    const double w = r2 - r0 + 1, h = r3 - r1 + 1, xc = r0 + 0.5 * (w - 1), yc = r1 + 0.5 * (h - 1);
    for (int k = 0; k < ns; ++k) {
      const double sw = w * (double)scales[k], sh = h * (double)scales[k];   // float32 scales as in the worker
      out[4 * n] = xc - 0.5 * (sw - 1);
      out[4 * n + 1] = yc - 0.5 * (sh - 1);
      out[4 * n + 2] = xc + 0.5 * (sw - 1);
      out[4 * n + 3] = yc + 0.5 * (sh - 1);
      ++n;
    }
  }
}

}  // namespace

extern "C" {

// gtmax_scratch: device uint64[B*G], zero-initialised by the caller.  Outputs follow the iterator's tensor
// contract (MNIteratorE2E.py:186-193): label [B,A*H*W] (a,h,w order, values -1/0/1), bbox_target / bbox_weight
// [B,4A,H,W].  argmax_out (optional) = matched GT index per anchor in flat (h,w,a) order, for parity tests.
int sniper_anchor_target(const float* gt_valid, const int* ngt, int G, const float* gt_invalid, const int* ninv, int Gi,
                         const float* im_info, const uint8_t* disable, int B, int H, int W, int feat_stride,
                         const float* scales, int ns, const float* ratios, int nr, double pos_thresh, double neg_thresh,
                         unsigned long long* gtmax_scratch, float* label, float* bbox_target, float* bbox_weight,
                         int32_t* argmax_out, void* stream) {
  SN_CHECK(ns * nr <= 64 && ns > 0 && nr > 0, "anchor_target: need 0 < ns*nr <= 64");
  AtArgs a;
  a.gt = gt_valid; a.ngt = ngt; a.inv = gt_invalid; a.ninv = ninv; a.im_info = im_info; a.disable = disable;
  a.B = B; a.G = G; a.Gi = Gi; a.H = H; a.W = W; a.A = ns * nr; a.stride = feat_stride;
  a.pos_thresh = pos_thresh; a.neg_thresh = neg_thresh;
  make_base_anchors(feat_stride, ratios, nr, scales, ns, a.base);
  a.gtmax = gtmax_scratch; a.label = label; a.target = bbox_target; a.weight = bbox_weight; a.argmax = argmax_out;
  const int total = H * W * a.A;
  dim3 grid(sn::div_up(total, 256), B);
  at_gtmax_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  SN_LAUNCH_CHECK();
  at_label_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  SN_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
