"""Seeded inputs shared by tests/test_refcuda_gpu.py (reference CUDA kernels vs oracle vs product on the B200) and
tests/golden/make_refcuda_golden.py (which records the reference kernels' outputs for the CPU suite)."""
import numpy as np

from sniper_b200 import synth

SCALES, RATIOS = (2, 4, 7, 10, 13, 16, 24), (0.5, 1, 2)


def mpt_case(seed, B, H=32, W=32, tie_fraction=0.0, dead_chip=False):
    cls_prob, bbox_pred, im_info, gts, vr = synth.mpt_inputs(seed, B, 21, H, W, tie_fraction=tie_fraction)
    if H != 32:
        im_info[:, 0], im_info[:, 1] = H * 16, W * 16
        gts[..., :4] = np.where(gts[..., 4:5] >= 0, np.clip(gts[..., :4], 0, H * 16 - 1), gts[..., :4])
    if dead_chip:
        vr[-1] = (5000.0, 6000.0)          # no proposal of the last chip is in range: every score -1 -> filler rows only
    return cls_prob, bbox_pred, im_info, gts, vr


def dpsroi_cases():
    """(name, data, rois, trans, kwargs) -- the reference's own unit-test shape (test_operator.py:4358-4389) and the
    ResNet-101 head's parameters (resnet_mx_101_e2e.py:286-293) on a reduced channel count."""
    rng = np.random.RandomState(3)
    data = rng.rand(2, 18, 14, 14).astype(np.float32)
    rois = np.array([[0, 10, 22, 161, 173], [1, 20, 15, 154, 160], [0, 0, 0, 5, 5], [1, 100, 100, 400, 300]], np.float32)
    trans = (rng.rand(4, 4, 3, 3).astype(np.float32) - 0.5)
    yield "unit", data, rois, trans, dict(spatial_scale=0.0625, output_dim=2, group_size=3, pooled=3, part_size=3, spp=4,
                                          trans_std=0.1)
    rng = np.random.RandomState(7)
    B, C = 2, 32
    data = rng.randn(B, C, 32, 32).astype(np.float32)
    rois = synth.rois_for_pool(rng, 24, B)
    rois[0, 1:] = (0, 0, 511, 511)
    rois[1, 1:] = (500, 500, 511, 511)
    rois[2, 1:] = (17.5, 33.25, 17.5, 33.25)
    trans = (rng.randn(24, 2, 7, 7) * 0.5).astype(np.float32)
    yield "head", data, rois, trans, dict(spatial_scale=0.0625, output_dim=C, group_size=1, pooled=7, part_size=7, spp=4,
                                          trans_std=0.1)


def psroi_case():
    rng = np.random.RandomState(9)
    data = rng.randn(2, 2 * 9, 20, 24).astype(np.float32)
    rois = synth.rois_for_pool(rng, 32, 2, chip=320)
    rois[0, 1:] = (5, 5, 5, 5)
    rois[1, 1:] = (300, 300, 319, 319)
    return data, rois, dict(spatial_scale=0.0625, output_dim=2, group_size=3, pooled=3)


def deform_case(seed=5, N=1, C=512, H=10, big_offsets=True):
    rng = np.random.RandomState(seed)
    x = rng.randn(N, C, H, H).astype(np.float32)
    off = (rng.randn(N, 72, H, H) * (1.5 if big_offsets else 0.05)).astype(np.float32)
    # a few offsets that land exactly on integer positions / on the border rows
    off[0, 0, 0, :4] = (2.0, -2.0, 1.0, 0.0)
    off[0, 1, 0, :4] = (2.0, 2.0, -1.0, 0.0)
    dcol = rng.randn(N, C * 9, H, H).astype(np.float32)
    return x, off, dcol
