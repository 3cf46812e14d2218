"""Post-processing of the multi-scale inference path (the half of lib/inference.py that is arithmetic rather than
bookkeeping): per-chip box decoding (`Tester.detect`, :100-139), score thresholding (`get_detections`, :291-301) and the
cross-scale aggregation with per-(image, class) soft-NMS (`Tester.aggregate`, :152-230).

The reference aggregates on the host and farms the NMS out to `Pool(32)` processes, one `cpu_soft_nms` call per image
and class; here every (image, class) problem of the run is ONE launch of `sniper_soft_nms_batched`.  `backend="host"`
runs the same code path through the C-ABI host `sniper_cpu_soft_nms` (bit-identical to the reference's Cython) -- the
parity tests compare the two.  Not built (needs OpenCV, absent from this image): FocusChip generation
(`lib/chips/chips_inference.py`: cv2.dilate + findContours).
"""
import numpy as np
import torch

from . import host, ops


def bbox_pred(boxes, box_deltas):
    """nonlinear_pred (lib/bbox/bbox_transform.py:93-133): class-agnostic rois [N,4] + deltas [N,4k] -> boxes [N,4k],
    float64 like the reference (`boxes.astype(np.float)`)."""
    if boxes.shape[0] == 0:
        return np.zeros((0, box_deltas.shape[1]))
    boxes = boxes.astype(float, copy=False)
    w = boxes[:, 2] - boxes[:, 0] + 1.0
    h = boxes[:, 3] - boxes[:, 1] + 1.0
    cx = boxes[:, 0] + 0.5 * (w - 1.0)
    cy = boxes[:, 1] + 0.5 * (h - 1.0)
    pcx = box_deltas[:, 0::4] * w[:, None] + cx[:, None]
    pcy = box_deltas[:, 1::4] * h[:, None] + cy[:, None]
    pw = np.exp(box_deltas[:, 2::4]) * w[:, None]
    ph = np.exp(box_deltas[:, 3::4]) * h[:, None]
    out = np.zeros(box_deltas.shape)
    out[:, 0::4] = pcx - 0.5 * (pw - 1.0)
    out[:, 1::4] = pcy - 0.5 * (ph - 1.0)
    out[:, 2::4] = pcx + 0.5 * (pw - 1.0)
    out[:, 3::4] = pcy + 0.5 * (ph - 1.0)
    return out


def clip_boxes(boxes, im_shape):
    boxes[:, 0::4] = np.maximum(np.minimum(boxes[:, 0::4], im_shape[1] - 1), 0)
    boxes[:, 1::4] = np.maximum(np.minimum(boxes[:, 1::4], im_shape[0] - 1), 0)
    boxes[:, 2::4] = np.maximum(np.minimum(boxes[:, 2::4], im_shape[1] - 1), 0)
    boxes[:, 3::4] = np.maximum(np.minimum(boxes[:, 3::4], im_shape[0] - 1), 0)
    return boxes


def detect_postprocess(rois, cls_prob, bbox_deltas, im_info, batch_size):
    """Tester.detect :113-131 for one device's outputs: rois [B*R,5] (batch_idx, x1..y2), cls_prob [B*R,K], bbox_deltas
    [B*R,4], im_info [B,3] -> per chip (scores [R,K], boxes [R,4] in image scale)."""
    rois, cls_prob = np.asarray(rois), np.asarray(cls_prob)
    bbox_deltas, im_info = np.asarray(bbox_deltas), np.asarray(im_info)
    scores, preds = [], []
    for idx in range(batch_size):
        cids = np.where(rois[:, 0] == idx)[0]
        cboxes = clip_boxes(bbox_pred(rois[cids, 1:], bbox_deltas[cids]), im_info[idx, :2])
        preds.append(cboxes / im_info[idx, 2])
        scores.append(cls_prob[cids])
    return scores, preds


def threshold_detections(cscores, cboxes, num_classes, cls_thresh=1e-3, prune=None):
    """get_detections :291-301: per class j >= 1 the rows with score > cls_thresh as [n,5] (x1,y1,x2,y2,score).
    prune = (chip [x1,y1,x2,y2], im_width, im_height): additionally the `do_pruning` step (:335-351) -- rows moved to image
    coordinates and detections cut by a chip border dropped (chips_inference.project_and_prune), on all classes at once."""
    # one pass over the [R, K] score matrix instead of K np.where calls: (class, roi) pairs in class-major, roi-minor
    # order = the order the per-class loop of the reference produces
    cls_idx, roi_idx = np.nonzero(cscores[:, 1:num_classes].T > cls_thresh)
    rows = np.empty((len(roi_idx), 5), np.float32)
    rows[:, :4] = cboxes[roi_idx, 0:4]
    rows[:, 4] = cscores[roi_idx, cls_idx + 1]
    if prune is not None:
        from .chips_inference import prune_mask
        chip, im_w, im_h = prune
        rows = rows.astype(np.float64)
        rows[:, 0] += chip[0]; rows[:, 2] += chip[0]
        rows[:, 1] += chip[1]; rows[:, 3] += chip[1]
        keep = prune_mask(rows, chip, im_w, im_h)
        rows, cls_idx = rows[keep], cls_idx[keep]
    cuts = np.searchsorted(cls_idx, np.arange(1, num_classes - 1))
    return [np.zeros((0, 5), rows.dtype)] + np.split(rows, cuts)


def _valid_range_filter(cls_dets, valid_range):
    """Tester.aggregate :176-186 (areas = (x-extent) * (y-extent), no +1; lower bound exclusive, upper inclusive)."""
    heights = cls_dets[:, 2] - cls_dets[:, 0]
    widths = cls_dets[:, 3] - cls_dets[:, 1]
    areas = widths * heights
    lo = np.where(areas > valid_range[0] * valid_range[0])[0] if valid_range[0] > 0 else np.arange(len(areas))
    hi = np.where(areas <= valid_range[1] * valid_range[1])[0] if valid_range[1] > 0 else np.arange(len(areas))
    return cls_dets[np.intersect1d(lo, hi), :]


def aggregate(scale_cls_dets, valid_ranges, num_images, num_classes, sigma=0.55, nms_thresh=0.3, max_per_image=100,
              backend="device", device="cuda"):
    """Tester.aggregate (:152-230).  scale_cls_dets[s][j][i] = list over chips of [n,5] detections of class j in image i
    at scale s.  Returns all_boxes[j][i] = [n,5] after the valid-range filter, Gaussian soft-NMS (sigma, threshold
    0.001; nms_worker -> nms_wrapper -> cpu_soft_nms) and the MAX_PER_IMAGE cut."""
    assert len(scale_cls_dets) == len(valid_ranges), 'A valid range should be specified for each test scale'
    # Gather every detection of the run once, tagged with its (image, class) problem, apply each scale's valid range to
    # the whole block, and cut the (stably) sorted block into the per-problem arrays: rows of a problem keep the
    # (scale, chip, row) order in which the reference's nested loops stack them.
    blocks, keys = [], []
    for all_cls_dets, vr in zip(scale_cls_dets, valid_ranges):
        sb, sk = [], []
        for j in range(1, num_classes):
            for i in range(num_images):
                for cls_dets in all_cls_dets[j][i]:
                    d = np.asarray(cls_dets, np.float32).reshape(-1, 5)
                    if d.shape[0]:
                        sb.append(d)
                        sk.append(np.full(d.shape[0], i * (num_classes - 1) + (j - 1), np.int64))
        if sb:
            d, k = np.concatenate(sb, 0), np.concatenate(sk)
            areas = (d[:, 3] - d[:, 1]) * (d[:, 2] - d[:, 0])          # as _valid_range_filter (no +1)
            keep = np.ones(len(d), bool)
            if vr[0] > 0:
                keep &= areas > vr[0] * vr[0]
            if vr[1] > 0:
                keep &= areas <= vr[1] * vr[1]
            blocks.append(d[keep])
            keys.append(k[keep])
    nprob = num_images * (num_classes - 1)
    if blocks:
        d, k = np.concatenate(blocks, 0), np.concatenate(keys)
        order = np.argsort(k, kind="stable")
        d, k = np.ascontiguousarray(d[order]), k[order]
        cuts = np.searchsorted(k, np.arange(1, nprob))
        problems = np.split(d, cuts)
    else:
        problems = [np.zeros((0, 5), np.float32) for _ in range(nprob)]
    if backend == "device":
        sizes = [len(p) for p in problems]
        offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        flat = np.concatenate(problems, 0) if sum(sizes) else np.zeros((0, 5), np.float32)
        if len(flat):
            out, counts = ops.soft_nms_batched(torch.from_numpy(flat).to(device), torch.from_numpy(offsets).to(device),
                                               sigma=sigma, Nt=nms_thresh, threshold=0.001, method=2)
            out, counts = out.cpu().numpy(), counts.cpu().numpy()
        else:
            out, counts = flat, np.zeros(len(problems), np.int32)
        final = [out[offsets[p]:offsets[p] + counts[p]] for p in range(len(problems))]
    else:
        final = [host.cpu_soft_nms(p.copy(), sigma=sigma, Nt=nms_thresh, threshold=0.001, method=2) if len(p) else p
                 for p in problems]
    all_boxes = [[[] for _ in range(num_images)] for _ in range(num_classes)]
    for i in range(num_images):
        for j in range(1, num_classes):
            all_boxes[j][i] = final[i * (num_classes - 1) + (j - 1)]
    for i in range(num_images):
        if max_per_image > 0:
            image_scores = np.hstack([all_boxes[j][i][:, -1] for j in range(1, num_classes)])
            if len(image_scores) > max_per_image:
                image_thresh = np.sort(image_scores)[-max_per_image]
                for j in range(1, num_classes):
                    keep = np.where(all_boxes[j][i][:, -1] >= image_thresh)[0]
                    all_boxes[j][i] = all_boxes[j][i][keep, :]
    return all_boxes
