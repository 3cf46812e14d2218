// TEST INFRASTRUCTURE ONLY (oracle/).  C harness around the reference's OWN CPU operators, compiled from the sources
// where they lie (/root/reference/SNIPER-mxnet/src/operator/multi_proposal_target.cc and multi_proposal.cc, see
// oracle/Makefile) into oracle/_ref/libref_mpt.so.  Nothing of the reference is copied: this file only instantiates
// the reference's OperatorProperty through its public interface (Init -> CreateOperator -> Forward on TBlobs that
// wrap caller buffers), exactly what MXNet's executor does (include/mxnet/operator.h).
#include <mxnet/operator.h>
#include <dmlc/registry.h>
#include <string>
#include <utility>
#include <vector>
#ifdef REF_TARGET
#include "multi_proposal_target-inl.h"
#else
#include "multi_proposal-inl.h"
#endif

// the one symbol of libmxnet the two translation units need
namespace dmlc {
DMLC_REGISTRY_ENABLE(::mxnet::OperatorPropertyReg);
}

namespace {

using mxnet::TBlob;
using mxnet::TShape;

TBlob blob(float* p, std::initializer_list<mxnet::index_t> dims) {
  TShape s(dims.begin(), dims.end());
  return TBlob(p, s, mshadow::cpu::kDevMask);
}

std::string tuple_str(const float* v, int n) {
  std::string s = "(";
  for (int i = 0; i < n; ++i) s += std::to_string(v[i]) + (i + 1 < n ? "," : "");
  return s + ")";
}

}  // namespace

extern "C" {

#ifdef REF_TARGET
// MultiProposalTargetOp<cpu>::Forward (multi_proposal_target.cc:237-497).
// cls_prob [B,2A,H,W], bbox_pred [B,4A,H,W], im_info [B,3], gt_boxes [B,100,5], valid_ranges [B,2]  ->
// rois [B*post,5], label [B*post,1], bbox_target [B*post,4], bbox_weight [B*post,4].  Returns 0, -1 on exception.
int ref_multi_proposal_target(float* cls_prob, float* bbox_pred, float* im_info, float* gt_boxes, float* valid_ranges,
                              int B, int A, int H, int W, int post_nms_top_n, int feat_stride, const float* scales,
                              int ns, const float* ratios, int nr, float threshold, float bbox_scale, float* rois,
                              float* label, float* bbox_target, float* bbox_weight) {
  try {
    mxnet::op::MultiProposalTargetProp prop;
    std::vector<std::pair<std::string, std::string> > kw = {
        {"rpn_post_nms_top_n", std::to_string(post_nms_top_n)}, {"threshold", std::to_string(threshold)},
        {"batch_size", std::to_string(B)},                      {"bbox_scale", std::to_string(bbox_scale)},
        {"feature_stride", std::to_string(feat_stride)},        {"scales", tuple_str(scales, ns)},
        {"ratios", tuple_str(ratios, nr)}};
    prop.Init(kw);
    mxnet::Operator* op = prop.CreateOperator(mxnet::Context::CPU());
    mxnet::OpContext ctx;
    ctx.is_train = 1;
    ctx.run_ctx.ctx = mxnet::Context::CPU();
    ctx.run_ctx.stream = nullptr;
    const mxnet::index_t n = (mxnet::index_t)B * post_nms_top_n;
    // cls_prob is handed over as (B, 2, A*H, W) -- the same memory as (B, 2A, H, W): the CPU operator reads the
    // foreground score as scores[b][1][a*H + h][w] (multi_proposal_target.cc:308), i.e. it assumes that view.
    std::vector<TBlob> in = {blob(cls_prob, {(mxnet::index_t)B, 2, (mxnet::index_t)(A * H), (mxnet::index_t)W}),
                             blob(bbox_pred, {(mxnet::index_t)B, (mxnet::index_t)(4 * A), (mxnet::index_t)H, (mxnet::index_t)W}),
                             blob(im_info, {(mxnet::index_t)B, 3}), blob(gt_boxes, {(mxnet::index_t)B, 100, 5}),
                             blob(valid_ranges, {(mxnet::index_t)B, 2})};
    std::vector<TBlob> out = {blob(rois, {n, 5}), blob(label, {n, 1}), blob(bbox_target, {n, 4}),
                              blob(bbox_weight, {n, 4})};
    std::vector<mxnet::OpReqType> req(4, mxnet::kWriteTo);
    op->Forward(ctx, in, req, out, {});
    delete op;
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "ref_multi_proposal_target: %s\n", e.what());
    return -1;
  }
}

#else
// MultiProposalGPUOp<cpu>::Forward (multi_proposal.cc:273-374) -- the inference proposal operator.
// cls_prob [B,2A,H,W], bbox_pred [B,4A,H,W], im_info [B,3] -> rois [B*post,5], scores [B*post].  Rows past the kept
// count are rand() fillers (multi_proposal.cc:360-367).
int ref_multi_proposal(float* cls_prob, float* bbox_pred, float* im_info, int B, int A, int H, int W, int pre_nms_top_n,
                       int post_nms_top_n, int rpn_min_size, int feat_stride, const float* scales, int ns,
                       const float* ratios, int nr, float threshold, float* rois, float* scores) {
  try {
    mxnet::op::MultiProposalProp prop;
    std::vector<std::pair<std::string, std::string> > kw = {
        {"rpn_pre_nms_top_n", std::to_string(pre_nms_top_n)}, {"rpn_post_nms_top_n", std::to_string(post_nms_top_n)},
        {"threshold", std::to_string(threshold)},             {"rpn_min_size", std::to_string(rpn_min_size)},
        {"feature_stride", std::to_string(feat_stride)},      {"scales", tuple_str(scales, ns)},
        {"ratios", tuple_str(ratios, nr)},                    {"batch_size", std::to_string(B)}};
    prop.Init(kw);
    mxnet::Operator* op = prop.CreateOperator(mxnet::Context::CPU());
    mxnet::OpContext ctx;
    ctx.is_train = 0;
    ctx.run_ctx.ctx = mxnet::Context::CPU();
    ctx.run_ctx.stream = nullptr;
    const mxnet::index_t n = (mxnet::index_t)B * post_nms_top_n;
    std::vector<TBlob> in = {blob(cls_prob, {(mxnet::index_t)B, (mxnet::index_t)(2 * A), (mxnet::index_t)H, (mxnet::index_t)W}),
                             blob(bbox_pred, {(mxnet::index_t)B, (mxnet::index_t)(4 * A), (mxnet::index_t)H, (mxnet::index_t)W}),
                             blob(im_info, {(mxnet::index_t)B, 3})};
    std::vector<TBlob> out = {blob(rois, {n, 5}), blob(scores, {n})};
    std::vector<mxnet::OpReqType> req(2, mxnet::kWriteTo);
    op->Forward(ctx, in, req, out, {});
    delete op;
    return 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "ref_multi_proposal: %s\n", e.what());
    return -1;
  }
}

#endif

}  // extern "C"
