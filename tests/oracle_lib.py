"""ctypes access to oracle/liboracle.so (CPU oracle -- test infrastructure only)."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None
_REF = None

P, I, F, D = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"])
        _LIB = ctypes.CDLL(so)
        _LIB.oracle_expf.restype = F
        _LIB.oracle_expf.argtypes = [F]
    return _LIB


def ref_chips():
    """The reference's own lib/chips/cchips.cpp compiled into oracle/_ref (None if unavailable)."""
    global _REF
    if _REF is None:
        so = os.path.join(ORACLE_DIR, "_ref", "libref_chips.so")
        if not os.path.exists(so):
            return None
        _REF = ctypes.CDLL(so)
    return _REF


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def generate_anchors(feat_stride, scales, ratios):
    s, r = f32(scales), f32(ratios)
    out = np.zeros((len(r) * len(s), 4), np.float32)
    lib().oracle_generate_anchors(I(feat_stride), _p(r), I(len(r)), _p(s), I(len(s)), _p(out))
    return out


def multi_proposal_target(cls_prob, bbox_pred, im_info, gt_boxes, valid_ranges, feat_stride=16,
                          scales=(2, 4, 7, 10, 13, 16, 24), ratios=(0.5, 1, 2), post=300):
    """NCHW inputs. Returns dict(rois,label,bbox_target,bbox_weight,keep_idx,num_kept,dets)."""
    cls_prob, bbox_pred = f32(cls_prob), f32(bbox_pred)
    im_info, gt_boxes, valid_ranges = f32(im_info), f32(gt_boxes), f32(valid_ranges)
    B, A4, H, W = bbox_pred.shape
    A = A4 // 4
    s, r = f32(scales), f32(ratios)
    rois = np.zeros((B * post, 5), np.float32)
    label = np.zeros((B * post,), np.float32)
    bt = np.zeros((B * post, 4), np.float32)
    bw = np.zeros((B * post, 4), np.float32)
    keep = np.zeros((B * post,), np.int32)
    nk = np.zeros((B,), np.int32)
    dets = np.zeros((B * A * H * W, 6), np.float32)
    rc = lib().oracle_multi_proposal_target(_p(cls_prob), _p(bbox_pred), _p(im_info), _p(gt_boxes), _p(valid_ranges),
                                            I(B), I(A), I(H), I(W), I(gt_boxes.shape[1]), I(post), I(feat_stride),
                                            _p(s), I(len(s)), _p(r), I(len(r)), _p(rois), _p(label), _p(bt), _p(bw),
                                            _p(keep), _p(nk), _p(dets))
    assert rc == 0
    return dict(rois=rois, label=label, bbox_target=bt, bbox_weight=bw, keep_idx=keep, num_kept=nk, dets=dets)


def multi_proposal_target_cpuop(cls_prob, bbox_pred, im_info, gt_boxes, valid_ranges, feat_stride=16,
                                scales=(2, 4, 7, 10, 13, 16, 24), ratios=(0.5, 1, 2), post=300, bbox_scale=1.0):
    """oracle/mpt_cpuop.c: restatement of the reference CPU operator (multi_proposal_target.cc).  gt_boxes [B,100,5]."""
    cls_prob, bbox_pred = f32(cls_prob), f32(bbox_pred)
    im_info, gt_boxes, valid_ranges = f32(im_info), f32(gt_boxes), f32(valid_ranges)
    B, A4, H, W = bbox_pred.shape
    A = A4 // 4
    assert gt_boxes.shape[1] == 100
    s, r = f32(scales), f32(ratios)
    rois = np.zeros((B * post, 5), np.float32)
    label = np.zeros((B * post,), np.float32)
    bt = np.zeros((B * post, 4), np.float32)
    bw = np.zeros((B * post, 4), np.float32)
    keep = np.zeros((B * post,), np.int32)
    nk = np.zeros((B,), np.int32)
    dets = np.zeros((B * A * H * W, 5), np.float32)
    rc = lib().oracle_multi_proposal_target_cpuop(_p(cls_prob), _p(bbox_pred), _p(im_info), _p(gt_boxes), _p(valid_ranges),
                                                  I(B), I(A), I(H), I(W), I(post), I(feat_stride), _p(s), I(len(s)),
                                                  _p(r), I(len(r)), F(bbox_scale), _p(rois), _p(label), _p(bt), _p(bw),
                                                  _p(keep), _p(nk), _p(dets))
    assert rc == 0
    return dict(rois=rois, label=label, bbox_target=bt, bbox_weight=bw, keep_idx=keep, num_kept=nk, dets=dets)


def multi_proposal(cls_prob, bbox_pred, im_info, feat_stride=16, scales=(2, 4, 7, 10, 13, 16, 24), ratios=(0.5, 1, 2),
                   pre=12000, post=300, flags=0, roi_iou_thresh=0.3, libm_exp=False):
    """oracle/mp_cpuop.c: the inference proposal operator (multi_proposal.cc; flags 1 = anchor-type suppression,
    2 = FastNMS of multi_proposal.cu).  Returns dict(rois, scores, keep_idx, num_kept)."""
    cls_prob, bbox_pred, im_info = f32(cls_prob), f32(bbox_pred), f32(im_info)
    B, A4, H, W = bbox_pred.shape
    A = A4 // 4
    s, r = f32(scales), f32(ratios)
    rois = np.zeros((B * post, 5), np.float32)
    scores = np.zeros((B * post,), np.float32)
    keep = np.zeros((B * post,), np.int32)
    nk = np.zeros((B,), np.int32)
    rc = lib().oracle_multi_proposal(_p(cls_prob), _p(bbox_pred), _p(im_info), I(B), I(A), I(H), I(W), I(pre), I(post),
                                     I(feat_stride), _p(s), I(len(s)), _p(r), I(len(r)), I(flags), F(roi_iou_thresh),
                                     I(int(libm_exp)), _p(rois), _p(scores), _p(keep), _p(nk))
    assert rc == 0
    return dict(rois=rois, scores=scores, keep_idx=keep, num_kept=nk)


_REF_OPS = {}


def ref_op_lib(name):
    """oracle/_ref/libref_mpt.so / libref_mp.so: the reference's own CPU operators (multi_proposal_target.cc,
    multi_proposal.cc) compiled from /root/reference by oracle/Makefile.  None if not built."""
    if name not in _REF_OPS:
        so = os.path.join(ORACLE_DIR, "_ref", name)
        _REF_OPS[name] = ctypes.CDLL(so) if os.path.exists(so) else None
    return _REF_OPS[name]


def ref_multi_proposal_target(cls_prob, bbox_pred, im_info, gt_boxes, valid_ranges, feat_stride=16,
                              scales=(2, 4, 7, 10, 13, 16, 24), ratios=(0.5, 1, 2), post=300, bbox_scale=1.0,
                              threshold=0.7):
    """Runs MultiProposalTargetOp<cpu>::Forward of the REFERENCE binary.  Returns None if the binary is absent."""
    L = ref_op_lib("libref_mpt.so")
    if L is None:
        return None
    cls_prob, bbox_pred = f32(cls_prob).copy(), f32(bbox_pred).copy()
    im_info, gt_boxes, valid_ranges = f32(im_info).copy(), f32(gt_boxes).copy(), f32(valid_ranges).copy()
    B, A4, H, W = bbox_pred.shape
    A = A4 // 4
    s, r = f32(scales), f32(ratios)
    rois = np.zeros((B * post, 5), np.float32)
    label = np.zeros((B * post, 1), np.float32)
    bt = np.zeros((B * post, 4), np.float32)
    bw = np.zeros((B * post, 4), np.float32)
    rc = L.ref_multi_proposal_target(_p(cls_prob), _p(bbox_pred), _p(im_info), _p(gt_boxes), _p(valid_ranges), I(B), I(A),
                                     I(H), I(W), I(post), I(feat_stride), _p(s), I(len(s)), _p(r), I(len(r)),
                                     F(threshold), F(bbox_scale), _p(rois), _p(label), _p(bt), _p(bw))
    assert rc == 0
    return dict(rois=rois, label=label[:, 0], bbox_target=bt, bbox_weight=bw)


def ref_multi_proposal(cls_prob, bbox_pred, im_info, feat_stride=16, scales=(2, 4, 7, 10, 13, 16, 24),
                       ratios=(0.5, 1, 2), pre=12000, post=300, min_size=3, threshold=0.7):
    """Runs MultiProposalGPUOp<cpu>::Forward (multi_proposal.cc) of the REFERENCE binary: (rois, scores)."""
    L = ref_op_lib("libref_mp.so")
    if L is None:
        return None
    cls_prob, bbox_pred, im_info = f32(cls_prob).copy(), f32(bbox_pred).copy(), f32(im_info).copy()
    B, A4, H, W = bbox_pred.shape
    A = A4 // 4
    s, r = f32(scales), f32(ratios)
    rois = np.zeros((B * post, 5), np.float32)
    scores = np.zeros((B * post,), np.float32)
    rc = L.ref_multi_proposal(_p(cls_prob), _p(bbox_pred), _p(im_info), I(B), I(A), I(H), I(W), I(pre), I(post),
                              I(min_size), I(feat_stride), _p(s), I(len(s)), _p(r), I(len(r)), F(threshold), _p(rois),
                              _p(scores))
    assert rc == 0
    return rois, scores


def deform_psroi_fwd(data, rois, trans, spatial_scale, output_dim, group_size, pooled, part_size, spp, trans_std,
                     no_trans):
    data, rois = f32(data), f32(rois)
    N = rois.shape[0]
    B, C, H, W = data.shape
    ncls = 1 if no_trans else trans.shape[1] // 2
    trans = None if no_trans else f32(trans)
    out = np.zeros((N, output_dim, pooled, pooled), np.float32)
    cnt = np.zeros_like(out)
    sidx = np.zeros((out.size, spp * spp, 4), np.int32)
    lib().oracle_deform_psroi_fwd(_p(data), _p(rois), _p(trans), I(N), I(C), I(H), I(W), F(spatial_scale),
                                  I(output_dim), I(group_size), I(pooled), I(part_size or pooled), I(spp), F(trans_std),
                                  I(int(no_trans)), I(ncls), _p(out), _p(cnt), _p(sidx))
    return out, cnt, sidx


def deform_psroi_bwd(top_diff, top_count, data, rois, trans, spatial_scale, output_dim, group_size, pooled, part_size,
                     spp, trans_std, no_trans):
    data, rois, top_diff, top_count = f32(data), f32(rois), f32(top_diff), f32(top_count)
    N = rois.shape[0]
    B, C, H, W = data.shape
    ncls = 1 if no_trans else trans.shape[1] // 2
    trans = None if no_trans else f32(trans)
    dd = np.zeros(data.shape, np.float64)
    td = None if no_trans else np.zeros(trans.shape, np.float64)
    lib().oracle_deform_psroi_bwd(_p(top_diff), _p(top_count), _p(data), _p(rois), _p(trans), I(N), I(C), I(H), I(W),
                                  F(spatial_scale), I(output_dim), I(group_size), I(pooled), I(part_size or pooled),
                                  I(spp), F(trans_std), I(int(no_trans)), I(ncls), _p(dd), _p(td))
    return dd, td


def psroi_fwd(data, rois, spatial_scale, output_dim, group_size, pooled):
    data, rois = f32(data), f32(rois)
    N = rois.shape[0]
    B, C, H, W = data.shape
    out = np.zeros((N, output_dim, pooled, pooled), np.float32)
    bins = np.zeros((out.size, 4), np.int32)
    lib().oracle_psroi_fwd(_p(data), _p(rois), I(N), I(C), I(H), I(W), F(spatial_scale), I(output_dim), I(group_size),
                           I(pooled), _p(out), _p(bins))
    return out, bins


def psroi_bwd(top_diff, rois, data_shape, spatial_scale, output_dim, group_size, pooled):
    top_diff, rois = f32(top_diff), f32(rois)
    B, C, H, W = data_shape
    dd = np.zeros(data_shape, np.float64)
    lib().oracle_psroi_bwd(_p(top_diff), _p(rois), I(rois.shape[0]), I(C), I(H), I(W), F(spatial_scale), I(output_dim),
                           I(group_size), I(pooled), _p(dd))
    return dd


def cpu_nms(dets, thresh, order=None):
    dets = f32(dets)
    if order is None:
        order = dets[:, 4].argsort()[::-1]
    order = np.ascontiguousarray(order, dtype=np.int64)
    keep = np.zeros((dets.shape[0],), np.int32)
    lib().oracle_cpu_nms.restype = I
    n = lib().oracle_cpu_nms(_p(dets), _p(order), I(dets.shape[0]), D(thresh), _p(keep))
    return keep[:n].copy()


def cpu_soft_nms(boxes, sigma=0.5, Nt=0.3, threshold=0.001, method=2):
    b = f32(boxes).copy()
    lib().oracle_cpu_soft_nms.restype = I
    n = lib().oracle_cpu_soft_nms(_p(b), I(b.shape[0]), F(sigma), F(Nt), F(threshold), ctypes.c_uint(method))
    return b[:n].copy()


def bbox_overlaps(boxes, query, ignore=False):
    boxes = np.ascontiguousarray(boxes, np.float64)
    query = np.ascontiguousarray(query, np.float64)
    out = np.zeros((boxes.shape[0], query.shape[0]), np.float64)
    lib().oracle_bbox_overlaps(_p(boxes), I(boxes.shape[0]), _p(query), I(query.shape[0]), _p(out), I(int(ignore)))
    return out


def ref_chips_generate(boxes, width, height, chipsize, stride, seed=1):
    L = ref_chips()
    boxes = f32(boxes)
    out = np.zeros((4096, 4), np.float32)
    L.ref_chips_generate.restype = I
    n = L.ref_chips_generate(_p(boxes), I(boxes.shape[0]), I(width), I(height), I(chipsize), I(stride),
                             ctypes.c_uint(seed), _p(out), I(4096))
    return out[:n].copy()
