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


def threshold_detections(cscores, cboxes, num_classes, cls_thresh=1e-3):
    """get_detections :291-301: per class j >= 1 the rows with score > cls_thresh as [n,5] (x1,y1,x2,y2,score)."""
    out = [np.zeros((0, 5), np.float32)]
    for j in range(1, num_classes):
        inds = np.where(cscores[:, j] > cls_thresh)[0]
        out.append(np.hstack((cboxes[inds, 0:4], cscores[inds, j, np.newaxis])).astype(np.float32))
    return out


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
    problems = []
    for i in range(num_images):
        for j in range(1, num_classes):
            agg = np.empty((0, 5), dtype=np.float32)
            for all_cls_dets, vr in zip(scale_cls_dets, valid_ranges):
                for cls_dets in all_cls_dets[j][i]:
                    cls_dets = _valid_range_filter(np.asarray(cls_dets, np.float32).reshape(-1, 5), vr)
                    if cls_dets.shape[0] > 0:
                        agg = np.vstack((agg, cls_dets))
            problems.append(np.ascontiguousarray(agg, np.float32))
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
