"""ORACLE -- test infrastructure only.  Not part of the product path.

NumPy (Python 3) restatement of the reference's RPN anchor matching,
  lib/data_utils/data_workers.py  anchor_worker.__init__ :130-162 (anchor grid), worker :194-363
  lib/data_utils/generate_anchor.py:8-77 (np.round variant of the anchor table)
  lib/bbox/bbox.pyx:17-57 (float64 IoU, +1 convention), lib/bbox/bbox_transform.py:64-90 (nonlinear_transform)
with the RNG subsampling (npr.choice, :328-338) factored out into `subsample()` so that the pre-subsample
labels can be compared bit-exactly and the chosen disable indices fed to the device op.
Inputs are the already cropped/scaled/rounded/clipped/size-filtered boxes (the part of worker() before
:262 is dataset bookkeeping).  PARITY PIN: tests/test_oracle_cpu.py runs the reference's OWN anchor_worker.worker
(oracle/run_ref_anchor_worker.py executes the class from /root/reference with its generate_anchors, bbox_transform
and Cython bbox_overlaps; four Python-2 tokens rewritten in memory) and this file reproduces its labels (after the
reference's npr.choice subsampling), positive indices and targets bit for bit; fixture tests/golden/anchor_target_ref.npz.
"""
import numpy as np


def generate_anchors(base_size=16, ratios=(0.5, 1, 2), scales=(2, 4, 7, 10, 13, 16, 24)):
    ratios = np.array(ratios, dtype=np.float64)
    scales = np.array(scales, dtype=np.float32)

    def whctrs(a):
        w = a[2] - a[0] + 1
        h = a[3] - a[1] + 1
        return w, h, a[0] + 0.5 * (w - 1), a[1] + 0.5 * (h - 1)

    def mk(ws, hs, xc, yc):
        ws, hs = ws[:, None], hs[:, None]
        return np.hstack((xc - 0.5 * (ws - 1), yc - 0.5 * (hs - 1), xc + 0.5 * (ws - 1), yc + 0.5 * (hs - 1)))

    base = np.array([1, 1, base_size, base_size]) - 1
    w, h, xc, yc = whctrs(base)
    size_ratios = (w * h) / ratios
    ws = np.round(np.sqrt(size_ratios))
    hs = np.round(ws * ratios)
    ra = mk(ws, hs, xc, yc)
    out = []
    for i in range(ra.shape[0]):
        w, h, xc, yc = whctrs(ra[i])
        out.append(mk(np.array([w * s for s in scales]), np.array([h * s for s in scales]), xc, yc))
    return np.vstack(out)


def all_anchors(feat_h, feat_w, stride, ratios, scales):
    base = generate_anchors(stride, ratios, scales)
    A = base.shape[0]
    sx, sy = np.meshgrid(np.arange(0, feat_w) * stride, np.arange(0, feat_h) * stride)
    shifts = np.vstack((sx.ravel(), sy.ravel(), sx.ravel(), sy.ravel())).transpose()
    K = shifts.shape[0]
    return (base.reshape((1, A, 4)) + shifts.reshape((1, K, 4)).transpose((1, 0, 2))).reshape((K * A, 4)), A, K


def bbox_overlaps(boxes, query):
    N, K = boxes.shape[0], query.shape[0]
    ov = np.zeros((N, K), np.float64)
    for k in range(K):
        qa = (query[k, 2] - query[k, 0] + 1) * (query[k, 3] - query[k, 1] + 1)
        iw = np.minimum(boxes[:, 2], query[k, 2]) - np.maximum(boxes[:, 0], query[k, 0]) + 1
        ih = np.minimum(boxes[:, 3], query[k, 3]) - np.maximum(boxes[:, 1], query[k, 1]) + 1
        m = (iw > 0) & (ih > 0)
        ua = (boxes[:, 2] - boxes[:, 0] + 1) * (boxes[:, 3] - boxes[:, 1] + 1) + qa - iw * ih
        ov[m, k] = (iw * ih)[m] / ua[m]
    return ov


def nonlinear_transform(ex, gt):
    ew = ex[:, 2] - ex[:, 0] + 1.0
    eh = ex[:, 3] - ex[:, 1] + 1.0
    ecx = ex[:, 0] + 0.5 * (ew - 1.0)
    ecy = ex[:, 1] + 0.5 * (eh - 1.0)
    gw = gt[:, 2] - gt[:, 0] + 1.0
    gh = gt[:, 3] - gt[:, 1] + 1.0
    gcx = gt[:, 0] + 0.5 * (gw - 1.0)
    gcy = gt[:, 1] + 0.5 * (gh - 1.0)
    return np.vstack(((gcx - ecx) / (ew + 1e-7), (gcy - ecy) / (eh + 1e-7), np.log(gw / (ew + 1e-7)),
                      np.log(gh / (eh + 1e-7)))).transpose()


def anchor_target(gt_boxes, invalid_boxes, im_info, feat_h=32, feat_w=32, stride=16, ratios=(0.5, 1, 2),
                  scales=(2, 4, 7, 10, 13, 16, 24), pos_thresh=0.5, neg_thresh=0.4):
    """Returns dict(labels [A*H*W] in the flat (h,w,a) order BEFORE subsampling, argmax [A*H*W] (-1 outside),
    targets [H*W*A,4] float32)."""
    anchors, A, K = all_anchors(feat_h, feat_w, stride, ratios, scales)
    total = K * A
    inside = np.where((anchors[:, 0] >= -32) & (anchors[:, 1] >= -32) & (anchors[:, 2] < im_info[0] + 32) &
                      (anchors[:, 3] < im_info[1] + 32))[0]
    an = anchors[inside]
    labels = np.empty((len(inside),), np.float32)
    labels.fill(-1)
    gt_boxes = np.asarray(gt_boxes, np.float64).reshape(-1, 4)
    invalid_boxes = np.asarray(invalid_boxes, np.float64).reshape(-1, 4)
    argmax = np.zeros((len(inside),), np.int64)
    if gt_boxes.size > 0:
        ov = bbox_overlaps(an.astype(np.float64), gt_boxes)
        if len(invalid_boxes) > 0:
            ovn = bbox_overlaps(an.astype(np.float64), invalid_boxes)
            max_ovn = ovn[np.arange(len(inside)), ovn.argmax(axis=1)]
        argmax = ov.argmax(axis=1)
        max_ov = ov[np.arange(len(inside)), argmax]
        gt_argmax = ov.argmax(axis=0)
        gt_max = ov[gt_argmax, np.arange(ov.shape[1])]
        gt_argmax = np.where(ov == gt_max)[0]
        labels[max_ov < neg_thresh] = 0
        labels[gt_argmax] = 1
        labels[max_ov >= pos_thresh] = 1
        if len(invalid_boxes) > 0:
            labels[max_ovn > 0.3] = -1
    else:
        labels[:] = 0
        if len(invalid_boxes) > 0:
            ovn = bbox_overlaps(an.astype(np.float64), invalid_boxes)
            max_ovn = ovn[np.arange(len(inside)), ovn.argmax(axis=1)]
            labels[max_ovn > 0.3] = -1
    targets = np.zeros((len(inside), 4), np.float32)
    if gt_boxes.size > 0:
        targets[:] = nonlinear_transform(an, gt_boxes[argmax, :4])
    lab_full = np.full((total,), -1, np.float32)
    lab_full[inside] = labels
    tg_full = np.zeros((total, 4), np.float32)
    tg_full[inside] = targets
    am_full = np.full((total,), -1, np.int64)
    am_full[inside] = argmax
    return dict(labels=lab_full, targets=tg_full, argmax=am_full, A=A, K=K)


def subsample(labels, rng, num_fg=128, batch_size=256):
    """data_workers.py:326-338 with an injected RandomState; returns the disable mask (flat (h,w,a) order)."""
    labels = labels.copy()
    disable = np.zeros(labels.shape, bool)
    fg = np.where(labels == 1)[0]
    if len(fg) > num_fg:
        d = rng.choice(fg, size=(len(fg) - num_fg), replace=False)
        labels[d] = -1
        disable[d] = True
    num_bg = batch_size - np.sum(labels == 1)
    bg = np.where(labels == 0)[0]
    if len(bg) > num_bg:
        d = rng.choice(bg, size=(len(bg) - num_bg), replace=False)
        labels[d] = -1
        disable[d] = True
    return disable


def pack(labels_hwa, targets_hwa, feat_h, feat_w, A):
    """data_workers.py:346-356: (h,w,a) flat -> label [A*H*W] (a,h,w), targets/weights [4A,H,W]."""
    lab = labels_hwa.reshape((1, feat_h, feat_w, A)).transpose(0, 3, 1, 2).reshape(A * feat_h * feat_w)
    lab = lab.astype(np.float16).astype(np.float32)
    tg = targets_hwa.reshape((feat_h, feat_w, A * 4)).transpose(2, 0, 1)
    w = np.zeros((labels_hwa.shape[0], 4), np.float32)
    w[labels_hwa == 1, :] = 1.0
    w = w.reshape((feat_h, feat_w, A * 4)).transpose(2, 0, 1)
    return lab, np.ascontiguousarray(tg * (w == 1)), np.ascontiguousarray(w)
