"""Seeded synthetic inputs of the shapes the SNIPER training path sees."""
import numpy as np

SCALES_RES101 = (2, 4, 7, 10, 13, 16, 24)
RATIOS = (0.5, 1, 2)


def gt_boxes(rng, B, max_gt=100, chip=512, n_lo=1, n_hi=20, num_classes=81):
    """[B,max_gt,5] (x1,y1,x2,y2,cls), pad rows = -1; sqrt(area) in [10,300] px (chip coordinates)."""
    out = -np.ones((B, max_gt, 5), np.float32)
    for b in range(B):
        n = int(rng.randint(n_lo, n_hi + 1))
        s = np.exp(rng.uniform(np.log(10), np.log(300), n))
        ar = np.exp(rng.uniform(np.log(0.5), np.log(2.0), n))
        w, h = s * np.sqrt(ar), s / np.sqrt(ar)
        cx, cy = rng.uniform(0, chip, n), rng.uniform(0, chip, n)
        x1, y1 = np.clip(cx - w / 2, 0, chip - 1), np.clip(cy - h / 2, 0, chip - 1)
        x2, y2 = np.clip(cx + w / 2, 0, chip - 1), np.clip(cy + h / 2, 0, chip - 1)
        out[b, :n, 0], out[b, :n, 1], out[b, :n, 2], out[b, :n, 3] = np.round(x1), np.round(y1), np.round(x2), np.round(y2)
        out[b, :n, 4] = rng.randint(1, num_classes, n)
    return out


def chip_meta(B, chip=512):
    """im_info [B,3] and valid_ranges [B,2] cycling through the three SNIPER scales (yml:76-101)."""
    scales = (3.0, 1.667, 0.8)
    ranges = ((0.0, 80.0), (32.0, 150.0), (120.0, -1.0))
    im_info = np.zeros((B, 3), np.float32)
    vr = np.zeros((B, 2), np.float32)
    for b in range(B):
        s = scales[b % 3]
        lo, hi = ranges[b % 3]
        im_info[b] = (chip, chip, s)
        vr[b] = (0 if lo < 0 else lo * s, chip if hi < 0 else hi * s)
    return im_info, vr


def rpn_outputs(rng, B, A=21, H=32, W=32, tie_fraction=0.0):
    """cls_prob [B,2A,H,W] (softmax pairs) and bbox_pred [B,4A,H,W] in the reference's NCHW layout."""
    logits = rng.randn(B, 2, A * H * W).astype(np.float32) * 2.0
    if tie_fraction > 0:
        # quantise some logits so that exact score ties occur (exercises the reference's tie order)
        m = rng.rand(B, 1, A * H * W) < tie_fraction
        logits = np.where(m, np.round(logits), logits).astype(np.float32)
    e = np.exp(logits - logits.max(1, keepdims=True))
    prob = (e / e.sum(1, keepdims=True)).astype(np.float32)
    cls_prob = prob.reshape(B, 2 * A, H, W)
    bbox_pred = (rng.randn(B, 4 * A, H, W) * 0.3).astype(np.float32)
    return cls_prob, bbox_pred


def mpt_inputs(seed, B, A=21, H=32, W=32, max_gt=100, tie_fraction=0.0):
    rng = np.random.RandomState(seed)
    cls_prob, bbox_pred = rpn_outputs(rng, B, A, H, W, tie_fraction)
    im_info, vr = chip_meta(B)
    gts = gt_boxes(rng, B, max_gt)
    return cls_prob, bbox_pred, im_info, gts, vr


def rois_for_pool(rng, N, B, chip=512):
    s = np.exp(rng.uniform(np.log(8), np.log(400), N))
    ar = np.exp(rng.uniform(np.log(0.5), np.log(2.0), N))
    w, h = s * np.sqrt(ar), s / np.sqrt(ar)
    cx, cy = rng.uniform(0, chip, N), rng.uniform(0, chip, N)
    rois = np.zeros((N, 5), np.float32)
    rois[:, 0] = rng.randint(0, B, N)
    rois[:, 1] = np.clip(cx - w / 2, 0, chip - 1)
    rois[:, 2] = np.clip(cy - h / 2, 0, chip - 1)
    rois[:, 3] = np.clip(cx + w / 2, 0, chip - 1)
    rois[:, 4] = np.clip(cy + h / 2, 0, chip - 1)
    return rois
