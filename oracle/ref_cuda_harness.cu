// ORACLE harness -- test infrastructure only (never linked into the product library).
//
// extern "C" launchers around the REFERENCE's own CUDA kernels, which oracle/build_ref_cuda.py cuts out of the reference
// tree at build time into oracle/_ref/cu/*.inc (nothing of the reference is stored in this repository).  This file
// only supplies (a) the handful of names the cut text expects from mshadow / nnvm / mxnet, (b) launchers that take raw
// device pointers, use the reference's own launch geometry where it matters for the result (the NMS kernel's
// 1024-thread argmax tree) and synchronise, so that tests can call them through ctypes with torch data_ptr()s.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <cmath>
#include <set>
#include <vector>

#define NUM_THREADS_NMS 1024
#define CUDA_KERNEL_LOOP(i, n) \
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += blockDim.x * gridDim.x)

// ---- names the cut text uses -----------------------------------------------------------------------------------
namespace nnvm {
template <typename T>
struct Tuple {   // nnvm::Tuple<float>: only ndim() and operator[] are used by GenerateAnchors
  std::vector<T> v;
  size_t ndim() const { return v.size(); }
  const T& operator[](size_t i) const { return v[i]; }
};
}  // namespace nnvm
enum OpReqType { kNullOp, kWriteTo, kWriteInplace, kAddTo };   // include/mxnet/op_attr_types.h:45-58

namespace refmpt {
namespace utils {
#include "mpt_anchors.inc"
#include "mpt_kernels.inc"
}  // namespace utils
}  // namespace refmpt
namespace refdps {
#include "dpsroi_fwd.inc"
#include "dpsroi_bwd.inc"
}  // namespace refdps
namespace refps {
#include "psroi_fwd.inc"
#include "psroi_bwd.inc"
}  // namespace refps
namespace refdc {
#include "dim2col_fwd.inc"
#include "dim2col_col2im.inc"
#include "dim2col_coord.inc"
}  // namespace refdc

namespace {
const int kBaseThreadNum = 256;   // mshadow::cuda::kBaseThreadNum
int num_blocks(long n) {          // mxnet_op::cuda_get_num_blocks
  long b = (n + kBaseThreadNum - 1) / kBaseThreadNum;
  return (int)(b > 65535 ? 65535 : (b < 1 ? 1 : b));
}
int done(const char* what) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    fprintf(stderr, "ref_cuda_harness: %s: %s\n", what, cudaGetErrorString(e));
    return -1;
  }
  return 0;
}
__global__ void expf_kernel(const float* x, float* y, int n) {
  CUDA_KERNEL_LOOP(i, n) { y[i] = exp(x[i]); }   // the same overload getProps' `exp(dw)` resolves to (float)
}
}  // namespace

extern "C" {

// utils::GenerateAnchors as MultiProposalTargetGPUOp::Forward calls it (multi_proposal_target.cu:401-411); host.
int ref_generate_anchors(int feature_stride, const float* ratios, int nr, const float* scales, int ns, float* out) {
  std::vector<float> base_anchor(4);
  base_anchor[0] = 0.0;
  base_anchor[1] = 0.0;
  base_anchor[2] = feature_stride - 1.0;
  base_anchor[3] = feature_stride - 1.0;
  nnvm::Tuple<float> r, s;
  r.v.assign(ratios, ratios + nr);
  s.v.assign(scales, scales + ns);
  std::vector<float> anchors;
  refmpt::utils::GenerateAnchors(base_anchor, r, s, &anchors);
  for (size_t i = 0; i < anchors.size(); ++i) out[i] = anchors[i];
  return (int)anchors.size();
}

// getProps with the launch of multi_proposal_target.cu:416-419.  All pointers: device.
int ref_get_props(float* boxes, float* deltas, float* im_info, float* anchorbuf, float* scores, float* valid_ranges,
                  int num_images, int anchors, int height, int width, int stride) {
  const int total = num_images * anchors * height * width;
  const int threadsPerBlock = NUM_THREADS_NMS;
  const int numblocks = (total / threadsPerBlock) + 1;
  refmpt::utils::getProps<<<numblocks, threadsPerBlock>>>(boxes, deltas, im_info, anchorbuf, scores, valid_ranges,
                                                          num_images, anchors, height, width, stride);
  return done("getProps");
}

// NonMaximumSuppression with the launch of multi_proposal_target.cu:422 (one 1024-thread block per image).
int ref_nms(float* dets, int post_nms_top_n, int num_images, int num_anchors, int width, int height, float* propsout) {
  refmpt::utils::NonMaximumSuppression<<<num_images, NUM_THREADS_NMS>>>(dets, post_nms_top_n, num_images, num_anchors,
                                                                        width, height, propsout);
  return done("NonMaximumSuppression");
}

// The host half of MultiProposalTargetGPUOp::Forward (multi_proposal_target.cu:435-578), verbatim: GT append, IoU,
// labels, regression targets.  All pointers: HOST (the reference keeps them in host staging buffers); rois = the NMS
// kernel's propsout copied to the host.
void ref_mpt_host_assign(float* gt_boxes, float* rois, float* labels, float* bbox_targets, float* bbox_weights,
                         float* valid_ranges, int num_images, int rpn_post_nms_top_n) {
#include "mpt_host_assign.inc"
}

int ref_expf(const float* x, float* y, int n) {
  expf_kernel<<<num_blocks(n), kBaseThreadNum>>>(x, y, n);
  return done("expf");
}

// DeformablePSROIPoolForward / BackwardAcc launches (deformable_psroi_pooling.cu:164-199, 334-375)
int ref_dpsroi_fwd(int count, const float* bottom_data, float spatial_scale, int channels, int height, int width,
                   int pooled_height, int pooled_width, const float* bottom_rois, const float* bottom_trans,
                   int no_trans, float trans_std, int sample_per_part, int output_dim, int group_size, int part_size,
                   int num_classes, int channels_each_class, float* top_data, float* top_count) {
  refdps::DeformablePSROIPoolForwardKernel<float><<<num_blocks(count), kBaseThreadNum>>>(
      count, bottom_data, spatial_scale, channels, height, width, pooled_height, pooled_width, bottom_rois, bottom_trans,
      no_trans != 0, trans_std, sample_per_part, output_dim, group_size, part_size, num_classes, channels_each_class,
      top_data, top_count);
  return done("DeformablePSROIPoolForwardKernel");
}

int ref_dpsroi_bwd(int count, const float* top_diff, const float* top_count, int num_rois, float spatial_scale,
                   int channels, int height, int width, int pooled_height, int pooled_width, int output_dim,
                   float* bottom_data_diff, float* bottom_trans_diff, const float* bottom_data, const float* bottom_rois,
                   const float* bottom_trans, int no_trans, float trans_std, int sample_per_part, int group_size,
                   int part_size, int num_classes, int channels_each_class) {
  refdps::DeformablePSROIPoolBackwardAccKernel<float><<<num_blocks(count), kBaseThreadNum>>>(
      count, top_diff, top_count, num_rois, spatial_scale, channels, height, width, pooled_height, pooled_width,
      output_dim, bottom_data_diff, bottom_trans_diff, bottom_data, bottom_rois, bottom_trans, no_trans != 0, trans_std,
      sample_per_part, group_size, part_size, num_classes, channels_each_class);
  return done("DeformablePSROIPoolBackwardAccKernel");
}

// PSROIPoolForward / BackwardAcc launches (psroi_pooling.cu:121-142, 216-239)
int ref_psroi_fwd(int count, const float* bottom_data, float spatial_scale, int channels, int height, int width,
                  int pooled_height, int pooled_width, const float* bottom_rois, int output_dim, int group_size,
                  float* top_data) {
  refps::PSROIPoolForwardKernel<float><<<num_blocks(count), kBaseThreadNum>>>(
      count, bottom_data, spatial_scale, channels, height, width, pooled_height, pooled_width, bottom_rois, output_dim,
      group_size, top_data);
  return done("PSROIPoolForwardKernel");
}

int ref_psroi_bwd(int count, const float* top_diff, int num_rois, float spatial_scale, int channels, int height,
                  int width, int pooled_height, int pooled_width, int group_size, int output_dim, float* bottom_diff,
                  const float* bottom_rois) {
  refps::PSROIPoolBackwardAccKernel<float><<<num_blocks(count), kBaseThreadNum>>>(
      count, top_diff, num_rois, spatial_scale, channels, height, width, pooled_height, pooled_width, group_size,
      output_dim, bottom_diff, bottom_rois);
  return done("PSROIPoolBackwardAccKernel");
}

// deformable_im2col / col2im / col2im_coord for ONE image, as DeformableConvolutionOp loops over the batch
// (deformable_convolution-inl.h:140-160, 211-252).  data_im [C,H,W], data_offset [dg*2*kh*kw, Hc, Wc],
// data_col [C*kh*kw, Hc, Wc].
int ref_deform_im2col(const float* data_im, const float* data_offset, int channels, int height, int width, int kh,
                      int kw, int pad, int stride, int dil, int dgroups, int height_col, int width_col,
                      float* data_col) {
  const int n = channels * height_col * width_col;
  refdc::deformable_im2col_gpu_kernel<float><<<num_blocks(n), kBaseThreadNum>>>(
      n, data_im, data_offset, height, width, kh, kw, pad, pad, stride, stride, dil, dil, channels / dgroups, height_col,
      width_col, data_col);
  return done("deformable_im2col_gpu_kernel");
}

int ref_deform_col2im(const float* data_col, const float* data_offset, int channels, int height, int width, int kh,
                      int kw, int pad, int stride, int dil, int dgroups, int height_col, int width_col,
                      float* grad_im) {
  const int n = channels * kh * kw * height_col * width_col;
  refdc::deformable_col2im_gpu_kernel<float><<<num_blocks(n), kBaseThreadNum>>>(
      n, data_col, data_offset, channels, height, width, kh, kw, pad, pad, stride, stride, dil, dil, channels / dgroups,
      height_col, width_col, grad_im, kWriteTo);
  return done("deformable_col2im_gpu_kernel");
}

int ref_deform_col2im_coord(const float* data_col, const float* data_im, const float* data_offset, int channels,
                            int height, int width, int kh, int kw, int pad, int stride, int dil, int dgroups,
                            int height_col, int width_col, float* grad_offset) {
  const int n = height_col * width_col * 2 * kh * kw * dgroups;
  // NB: this kernel's `channel_per_deformable_group` counts COLUMN channels, col_shape[0] / deformable_group
  // (deformable_im2col.cuh:499), unlike the other two kernels (im_shape[1] / deformable_group)
  refdc::deformable_col2im_coord_gpu_kernel<float><<<num_blocks(n), kBaseThreadNum>>>(
      n, data_col, data_im, data_offset, channels, height, width, kh, kw, pad, pad, stride, stride, dil, dil,
      channels * kh * kw / dgroups, height_col, width_col, grad_offset, kWriteTo);
  return done("deformable_col2im_coord_gpu_kernel");
}

}  // extern "C"
