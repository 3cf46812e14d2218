"""GPU parity: fused MultiProposalTarget (C-ABI) vs the CPU oracle -- bit-exact on rois/keep/labels."""
import numpy as np
import pytest

import oracle_lib as O
from sniper_b200 import synth

pytestmark = pytest.mark.gpu


def _run_gpu(cls_prob, bbox_pred, im_info, gts, vr, layout=0, A=21, scales=synth.SCALES_RES101):
    import torch
    from sniper_b200 import ops
    dev = "cuda:0"
    cp, bp = torch.from_numpy(cls_prob).to(dev), torch.from_numpy(bbox_pred).to(dev)
    if layout == 1:
        cp = cp.permute(0, 2, 3, 1).contiguous()
        bp = bp.permute(0, 2, 3, 1).contiguous()
    out = ops.multi_proposal_target(cp, bp, torch.from_numpy(im_info).to(dev), torch.from_numpy(gts).to(dev),
                                    torch.from_numpy(vr).to(dev), scales=scales, layout=layout, return_keep=True)
    torch.cuda.synchronize()
    return [o.cpu().numpy() for o in out]


def _compare(res, gpu):
    rois, label, bt, bw, keep, nk = gpu
    np.testing.assert_array_equal(nk, res["num_kept"])
    np.testing.assert_array_equal(keep, res["keep_idx"])
    assert rois.tobytes() == res["rois"].tobytes()          # bit-exact boxes incl. fillers + GT rows
    np.testing.assert_array_equal(label, res["label"])
    np.testing.assert_array_equal(bw, res["bbox_weight"])
    # targets contain log(): double log on both sides, allow 1 float ulp
    np.testing.assert_allclose(bt, res["bbox_target"], rtol=2e-7, atol=1e-7)
    assert np.mean(bt != res["bbox_target"]) < 1e-3


@pytest.mark.parametrize("seed,B,H,W,layout", [(11, 2, 16, 16, 0), (12, 4, 32, 32, 0), (13, 3, 32, 32, 1)])
def test_mpt_small(seed, B, H, W, layout):
    inp = synth.mpt_inputs(seed, B, 21, H, W)
    res = O.multi_proposal_target(*inp)
    _compare(res, _run_gpu(*inp, layout=layout))


def test_mpt_full_batch20():
    inp = synth.mpt_inputs(21, 20, 21, 32, 32)
    res = O.multi_proposal_target(*inp)
    _compare(res, _run_gpu(*inp))


def test_mpt_fast_and_sequential_nms_agree(monkeypatch):
    # the sorted/bit-mask fast path and the round-by-round emulation must give the same bits as the oracle
    inp = synth.mpt_inputs(61, 6, 21, 32, 32)
    res = O.multi_proposal_target(*inp)
    monkeypatch.setenv("SNIPER_NMS_FAST", "1")
    _compare(res, _run_gpu(*inp))
    monkeypatch.setenv("SNIPER_NMS_FAST", "0")
    _compare(res, _run_gpu(*inp))
    # few valid candidates (narrow valid range): fewer than 300 kept, fillers must match too
    cls_prob, bbox_pred, im_info, gts, vr = synth.mpt_inputs(62, 3, 21, 32, 32)
    vr[:] = (200.0, 230.0)
    res = O.multi_proposal_target(cls_prob, bbox_pred, im_info, gts, vr)
    assert res["num_kept"].max() < 300
    monkeypatch.setenv("SNIPER_NMS_FAST", "1")
    _compare(res, _run_gpu(cls_prob, bbox_pred, im_info, gts, vr))


def test_mpt_fast_path_is_taken_and_handles_ties(monkeypatch):
    """The fast path must really run (fallback flag 0) on ordinary inputs, on near-constant scores (the
    random-init regime of bench.py: thousands of scores within 1e-2 of 0.5, exact float ties among them) and on
    quantised scores with short tie runs; only runs of > 64 equal scores may fall back."""
    import torch
    from sniper_b200 import ops
    monkeypatch.setenv("SNIPER_NMS_FAST", "1")
    rng = np.random.RandomState(71)
    B, A, H, W = 4, 21, 32, 32
    cases = []
    cases.append(synth.mpt_inputs(72, B, A, H, W))
    cp, bp, im, gt, vr = synth.mpt_inputs(73, B, A, H, W)
    lg = (rng.randn(B, 2, A * H * W) * 0.01).astype(np.float32)      # random-init regime
    e = np.exp(lg - lg.max(1, keepdims=True))
    cases.append(((e / e.sum(1, keepdims=True)).astype(np.float32).reshape(B, 2 * A, H, W), bp, im, gt, vr))
    lg = np.round(rng.randn(B, 2, A * H * W) * 64).astype(np.float32) / 64   # many short tie runs
    e = np.exp(lg - lg.max(1, keepdims=True))
    cases.append(((e / e.sum(1, keepdims=True)).astype(np.float32).reshape(B, 2 * A, H, W), bp, im, gt, vr))
    for ci, inp in enumerate(cases):
        res = O.multi_proposal_target(*inp)
        t = [torch.from_numpy(a).cuda() for a in inp]
        out = ops.multi_proposal_target(*t, return_keep=True, return_fallback=True)
        got = [o.cpu().numpy() for o in out]
        _compare(res, got[:6])
        if ci < 2:
            assert got[6].sum() == 0, "fast path fell back on case %d: %s" % (ci, got[6])
    d = res["dets"][:A * H * W]
    v = d[d[:, 4] != -1, 4]
    assert len(np.unique(v)) < len(v)      # the last case really contains exact ties


def test_mpt_score_ties_follow_reference_scan_order():
    # quantised logits -> thousands of exactly equal scores; keep order must follow the reference's
    # strided 3-level argmax (multi_proposal_target.cu:139-176)
    inp = synth.mpt_inputs(31, 3, 21, 32, 32, tie_fraction=0.9)
    res = O.multi_proposal_target(*inp)
    d = res["dets"][:21 * 32 * 32]
    v = d[d[:, 4] != -1, 4]
    assert len(np.unique(v)) < len(v) // 2  # the case really has ties
    _compare(res, _run_gpu(*inp))


def test_mpt_edge_cases():
    # chip with no GT, chip whose GTs are all out of range, chip where nothing survives the filters
    cls_prob, bbox_pred, im_info, gts, vr = synth.mpt_inputs(41, 4, 21, 32, 32)
    gts[0, :, :] = -1
    vr[1] = (400.0, 512.0)
    vr[2] = (1000.0, 2000.0)   # every proposal filtered -> 300 filler rows
    res = O.multi_proposal_target(cls_prob, bbox_pred, im_info, gts, vr)
    assert res["num_kept"][2] == 0
    _compare(res, _run_gpu(cls_prob, bbox_pred, im_info, gts, vr))


def test_mpt_mobilenet_shape():
    # config 4 geometry: stride 32, 16x16, A = 5 scales x 3 ratios
    scales = (1, 2, 4, 8, 12)
    rng = np.random.RandomState(51)
    B, A, H, W = 5, 15, 16, 16
    cls_prob, bbox_pred = synth.rpn_outputs(rng, B, A, H, W)
    im_info, vr = synth.chip_meta(B)
    gts = synth.gt_boxes(rng, B)
    res = O.multi_proposal_target(cls_prob, bbox_pred, im_info, gts, vr, feat_stride=32, scales=scales)
    import torch
    from sniper_b200 import ops
    t = lambda a: torch.from_numpy(a).cuda()
    out = ops.multi_proposal_target(t(cls_prob), t(bbox_pred), t(im_info), t(gts), t(vr), feat_stride=32,
                                    scales=scales, return_keep=True)
    _compare(res, [o.cpu().numpy() for o in out])


def test_mpt_error_reporting():
    import torch
    from sniper_b200 import ops
    from sniper_b200._lib import SniperError
    inp = synth.mpt_inputs(1, 1, 21, 2, 2)  # 84 anchors < 300 -> must be refused, not UB as in the reference
    t = lambda a: torch.from_numpy(a).cuda()
    with pytest.raises(SniperError):
        ops.multi_proposal_target(*[t(a) for a in inp])


# ------------------------------------------------------------------------------------------------
# Inference proposal operator (MultiProposal): CUDA path vs oracle/mp_cpuop.c (itself bit-identical to the reference's
# CPU operator binary with libm exp, tests/test_oracle_cpu.py); here both sides use the correctly rounded exp.
# ------------------------------------------------------------------------------------------------
def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _tie_free_scores(cls_prob, seed):
    B, A2, H, W = cls_prob.shape
    A = A2 // 2
    n = B * A * H * W
    fg = ((np.random.RandomState(1000 + seed).permutation(n) + 1.0) / (n + 1.0)).astype(np.float32).reshape(B, A, H, W)
    out = cls_prob.copy()
    out[:, A:] = fg
    out[:, :A] = 1.0 - fg
    return out


@pytest.mark.parametrize("seed,B,HW,fast", [(31, 2, 32, "1"), (32, 3, 32, "0"), (33, 1, 20, "1"), (34, 20, 32, "1")])
def test_multi_proposal_inference_op_bit_exact(seed, B, HW, fast):
    """rois, scores, kept anchor indices and counts == oracle, bit for bit; HW=32: 21504 anchors > 12000 (the exact
    top-12000 selection + compaction runs), HW=20: 8400 anchors (no selection); fast = sorted bit-mask NMS / sequential
    emulation; NCHW and NHWC inputs."""
    import os
    import torch
    from sniper_b200 import ops
    cls_prob, bbox_pred, im_info, gts, vr = synth.mpt_inputs(seed, B, 21, HW, HW)
    cls_prob = _tie_free_scores(cls_prob, seed)
    ref = O.multi_proposal(cls_prob, bbox_pred, im_info, libm_exp=False)
    os.environ["SNIPER_NMS_FAST"] = fast
    try:
        rois, scores, keep, nk = ops.multi_proposal(_t(cls_prob), _t(bbox_pred), _t(im_info), return_keep=True)
        cn = _t(np.ascontiguousarray(cls_prob.transpose(0, 2, 3, 1)))
        bn = _t(np.ascontiguousarray(bbox_pred.transpose(0, 2, 3, 1)))
        rois2, scores2 = ops.multi_proposal(cn, bn, _t(im_info), layout=ops.NHWC)
    finally:
        os.environ.pop("SNIPER_NMS_FAST", None)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(nk.cpu().numpy(), ref["num_kept"])
    np.testing.assert_array_equal(keep.cpu().numpy(), ref["keep_idx"])
    assert rois.cpu().numpy().tobytes() == ref["rois"].tobytes()
    assert scores.cpu().numpy().tobytes() == ref["scores"].tobytes()
    assert rois2.cpu().numpy().tobytes() == ref["rois"].tobytes() and scores2.cpu().numpy().tobytes() == ref["scores"].tobytes()


def test_multi_proposal_anchor_type_suppression_and_operator_api():
    import torch
    from sniper_b200 import ops, operator_py
    cls_prob, bbox_pred, im_info, gts, vr = synth.mpt_inputs(41, 2)
    cls_prob = _tie_free_scores(cls_prob, 41)
    ref = O.multi_proposal(cls_prob, bbox_pred, im_info, flags=1, libm_exp=False)
    rois, scores, keep, nk = ops.multi_proposal(_t(cls_prob), _t(bbox_pred), _t(im_info), suppress_anchor_types=True,
                                                return_keep=True)
    np.testing.assert_array_equal(keep.cpu().numpy(), ref["keep_idx"])
    assert rois.cpu().numpy().tobytes() == ref["rois"].tobytes()
    plain = O.multi_proposal(cls_prob, bbox_pred, im_info, libm_exp=False)
    out, score = operator_py.MultiProposal(_t(cls_prob), _t(bbox_pred), _t(im_info), batch_size="2",
                                           rpn_post_nms_top_n="300", feature_stride="16")
    assert out.shape == (600, 5) and score.shape == (600, 1)
    assert out.cpu().numpy().tobytes() == plain["rois"].tobytes()
    assert score.cpu().numpy().ravel().tobytes() == plain["scores"].tobytes()


@pytest.mark.parametrize("seed,B,HW,fast,flags", [(51, 2, 32, "1", 2), (52, 2, 32, "0", 2), (53, 1, 20, "1", 3), (54, 3, 32, "1", 3)])
def test_multi_proposal_fast_nms_variant_bit_exact(seed, B, HW, fast, flags):
    """FastNMS of the reference's GPU build (pair tests restricted to the anchor-overlap map, multi_proposal.cu:267-387)
    +/- anchor-type suppression: rois, scores (incl. the -1 of rows marked after they were kept), keep indices ==
    oracle/mp_cpuop.c, for the bit-mask and the sequential kernels."""
    import os
    import torch
    from sniper_b200 import ops
    cls_prob, bbox_pred, im_info, gts, vr = synth.mpt_inputs(seed, B, 21, HW, HW)
    cls_prob = _tie_free_scores(cls_prob, seed)
    ref = O.multi_proposal(cls_prob, bbox_pred, im_info, flags=flags, roi_iou_thresh=0.3, libm_exp=False)
    os.environ["SNIPER_NMS_FAST"] = fast
    try:
        rois, scores, keep, nk = ops.multi_proposal(_t(cls_prob), _t(bbox_pred), _t(im_info), fast_nms=True,
                                                    suppress_anchor_types=bool(flags & 1), roi_iou_thresh=0.3,
                                                    return_keep=True)
    finally:
        os.environ.pop("SNIPER_NMS_FAST", None)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(nk.cpu().numpy(), ref["num_kept"])
    np.testing.assert_array_equal(keep.cpu().numpy(), ref["keep_idx"])
    assert rois.cpu().numpy().tobytes() == ref["rois"].tobytes()
    assert scores.cpu().numpy().tobytes() == ref["scores"].tobytes()
    plain = O.multi_proposal(cls_prob, bbox_pred, im_info, flags=flags & 1, libm_exp=False)
    assert not np.array_equal(plain["keep_idx"], ref["keep_idx"])          # the restriction does change the result
