"""RPN anchor matching: device op vs the numpy oracle (labels/argmax bit-exact, targets <= 1 ulp)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import anchor_target_np as AT  # noqa: E402
from sniper_b200 import synth  # noqa: E402


def test_oracle_anchor_table_and_overlaps_cpu():
    a = AT.generate_anchors(16, (0.5, 1, 2), (2, 4, 7, 10, 13, 16, 24))
    assert a.shape == (21, 4)
    np.testing.assert_array_equal(a[7], [-8., -8., 23., 23.])          # ratio 1, scale 2
    w, h = a[:, 2] - a[:, 0] + 1, a[:, 3] - a[:, 1] + 1
    np.testing.assert_array_equal(w[:7], 23 * np.array([2, 4, 7, 10, 13, 16, 24]))   # np.round(sqrt(512)) = 23
    np.testing.assert_array_equal(h[:7], 12 * np.array([2, 4, 7, 10, 13, 16, 24]))
    import oracle_lib as O
    rng = np.random.RandomState(0)
    b1 = synth.rois_for_pool(rng, 40, 1)[:, 1:].astype(np.float64)
    b2 = synth.rois_for_pool(rng, 9, 1)[:, 1:].astype(np.float64)
    assert AT.bbox_overlaps(b1, b2).tobytes() == O.bbox_overlaps(b1, b2).tobytes()
    res = AT.anchor_target(b2[:5], b2[5:], (512, 512, 1.0))
    lab = res["labels"]
    assert set(np.unique(lab)).issubset({-1.0, 0.0, 1.0}) and (lab == 1).sum() >= 5
    d = AT.subsample(lab, np.random.RandomState(1))
    l2 = lab.copy(); l2[d] = -1
    assert (l2 == 1).sum() <= 128 and ((l2 == 1).sum() + (l2 == 0).sum()) <= 256


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1])
def test_anchor_target_matches_oracle(seed):
    import torch
    from sniper_b200 import ops
    rng = np.random.RandomState(seed)
    B, G, Gi = 4, 100, 100
    gts = synth.gt_boxes(rng, B, n_lo=0, n_hi=12)
    gtv = np.zeros((B, G, 4), np.float32); gti = np.zeros((B, Gi, 4), np.float32)
    ngt = np.zeros(B, np.int32); ninv = np.zeros(B, np.int32)
    im_info = np.array([[512, 512, 1.0], [512, 512, 3.0], [400, 512, 1.667], [512, 384, 0.8]], np.float32)
    disable = np.zeros((B, 32 * 32 * 21), np.uint8)
    refs = []
    for b in range(B):
        n = int((gts[b, :, 4] != -1).sum())
        boxes = gts[b, :n, :4]
        split = n // 3 if b != 1 else n            # chip 1: no valid GT at all
        inv, val = boxes[:split], boxes[split:]
        if b == 3:
            inv = boxes[:0]
        gtv[b, :len(val)] = val; ngt[b] = len(val)
        gti[b, :len(inv)] = inv; ninv[b] = len(inv)
        r = AT.anchor_target(val, inv, im_info[b])
        dis = AT.subsample(r["labels"], np.random.RandomState(10 + b))
        disable[b] = dis
        lab = r["labels"].copy(); lab[dis] = -1
        refs.append((r, AT.pack(lab, r["targets"], 32, 32, 21)))
    t = lambda a: torch.from_numpy(a).cuda()
    label, bt, bw, am = ops.anchor_target(t(gtv), t(ngt), t(gti), t(ninv), t(im_info), disable=t(disable), want_argmax=True)
    label, bt, bw, am = [x.cpu().numpy() for x in (label, bt, bw, am)]
    for b in range(B):
        r, (lab, tg, w) = refs[b]
        np.testing.assert_array_equal(label[b], lab)                       # labels: bit-exact
        np.testing.assert_array_equal(bw[b], w)
        if ngt[b] > 0:
            inside = r["argmax"] >= 0
            np.testing.assert_array_equal(am[b][inside], r["argmax"][inside].astype(np.int32))
        np.testing.assert_allclose(bt[b], tg, rtol=2e-7, atol=1e-7)        # double log on both sides
    assert (label == 1).sum() > 0 and (label == 0).sum() > 0
