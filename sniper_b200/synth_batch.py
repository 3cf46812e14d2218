"""Synthetic SNIPER chip batches with the tensor contract of MNIteratorE2E._get_batch
(lib/iterators/MNIteratorE2E.py:175-219)."""
import numpy as np
import torch

from . import synth


def _anchor_labels(rng, B, A, H, W, gts):
    """Cheap stand-in RPN labels/targets with the iterator's value ranges (label in {-1,0,1}, <=256 sampled
    per chip, sparse bbox weights); the real anchor matcher is the anchor_target op."""
    label = -np.ones((B, A * H * W), np.float32)
    bt = np.zeros((B, 4 * A, H, W), np.float32)
    bw = np.zeros((B, 4 * A, H, W), np.float32)
    for b in range(B):
        idx = rng.choice(A * H * W, 256, replace=False)
        fg = idx[:rng.randint(8, 64)]
        label[b, idx] = 0
        label[b, fg] = 1
        a, hw = fg // (H * W), fg % (H * W)
        for j in range(4):
            bt[b].reshape(4 * A, H * W)[4 * a + j, hw] = rng.randn(len(fg)) * 0.3
            bw[b].reshape(4 * A, H * W)[4 * a + j, hw] = 1.0
    return label, bt, bw


def make_batch(B, seed=3, device="cuda", chip=512, A=21, stride=16, pinned=False):
    rng = np.random.RandomState(seed)
    H = W = chip // stride
    data = (rng.randn(B, 3, chip, chip) * 60).astype(np.float32)
    gts = synth.gt_boxes(rng, B)
    im_info, vr = synth.chip_meta(B, chip)
    label, bt, bw = _anchor_labels(rng, B, A, H, W, gts)
    host = dict(data=data, label=label, bbox_target=bt, bbox_weight=bw, gt_boxes=gts, valid_ranges=vr, im_info=im_info)
    out = {}
    for k, v in host.items():
        t = torch.from_numpy(v)
        if device == "cpu":
            out[k] = t.pin_memory() if pinned else t
        else:
            out[k] = t.to(device)
    return out
