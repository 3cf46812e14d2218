"""The reference's OWN CUDA kernels, compiled from the reference tree (oracle/build_ref_cuda.py) and run on this GPU,
against the C oracle and the product kernels.  Pins what the reference implements on the GPU only:
  * MultiProposalTarget: getProps (min-size `&&` rule without +1, area / valid-range rule), NonMaximumSuppression
    (+1-free areas in the IoU, the 3-level argmax tie order, the row swap, filler boxes)  [multi_proposal_target.cu:116-331]
  * DeformablePSROIPooling / PSROIPooling forward + backward  [deformable_psroi_pooling.cu:48-330, psroi_pooling.cu:50-212]
  * deformable im2col / col2im / col2im_coord  [deformable_im2col.cuh:77-473]
`libref_cuda.so` is built with -fmad=false (the C abstract machine the oracle restates): comparisons are bit-exact
except for (a) CUDA's expf (<= 2 ulp, not reproducible on a CPU; the oracle and the product use a correctly rounded exp)
-- rows whose two exponentials agree must be bit-identical, (b) float atomics in the backward kernels (order-dependent
rounding; tolerance stated per test).  `libref_cuda_fma.so` = nvcc defaults (FMA contraction on, what the reference's
own build produces): its deltas against the no-FMA build are measured and bounded below.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
import ref_cuda_lib as R  # noqa: E402
import refcuda_cases as CASES  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not R.available(), reason="oracle/_ref/libref_cuda.so not built (needs /root/reference)")]
A, STRIDE = 21, 16


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _oracle_stages(inp, H, W):
    """oracle decode (dets) and the rois of its NMS stage BEFORE the GT append."""
    cls_prob, bbox_pred, im_info, gts, vr = inp
    B = cls_prob.shape[0]
    res = O.multi_proposal_target(*inp)
    dets = res["dets"].copy()
    ids = np.tile(np.arange(A * H * W, dtype=np.int32), B)
    rois = np.zeros((B * 300, 5), np.float32)
    keep = np.zeros(B * 300, np.int32)
    nk = np.zeros(B, np.int32)
    O.lib().oracle_nms(O._p(dets), O._p(ids), O.I(300), O.I(B), O.I(A), O.I(W), O.I(H), O._p(rois), O._p(keep), O._p(nk))
    return res, rois, nk


def _exp_agree(bbox_pred, fma):
    """per anchor (b,a,h,w): CUDA expf == correctly rounded exp for both dw and dh."""
    B, _, H, W = bbox_pred.shape
    d = bbox_pred.reshape(B, A, 4, H, W)
    x = np.ascontiguousarray(np.stack([d[:, :, 2], d[:, :, 3]], 0).reshape(-1))
    y_ref = R.expf(_t(x), fma).cpu().numpy()
    y_orc = np.zeros_like(x)
    O.lib().oracle_expf_array(O._p(x), O._p(y_orc), O.ctypes.c_long(x.size))
    same = (y_ref == y_orc).reshape(2, -1)
    ulp = np.abs(y_ref.view(np.int32).astype(np.int64) - y_orc.view(np.int32).astype(np.int64)).max()
    return same[0] & same[1], int(ulp)


@pytest.mark.parametrize("seed,B,H,tie", [(1, 3, 32, 0.0), (2, 20, 32, 0.0), (4, 2, 8, 0.3)])
def test_getprops_kernel(seed, B, H, tie):
    inp = CASES.mpt_case(seed, B, H, H, tie)
    cls_prob, bbox_pred, im_info, gts, vr = inp
    res = O.multi_proposal_target(*inp)
    anchors = R.generate_anchors(STRIDE, CASES.SCALES, CASES.RATIOS)
    for fma in (False, True):
        boxes = R.get_props(_t(bbox_pred), _t(im_info), _t(anchors), _t(cls_prob), _t(vr), B, A, H, H, STRIDE, fma).cpu().numpy()
        agree, ulp = _exp_agree(bbox_pred, fma)
        same_rows = (boxes.view(np.int32) == res["dets"].view(np.int32)).all(1)
        print("fma=%d: expf agrees on %.2f%% of anchors (max %d ulp); identical rows %.2f%%"
              % (fma, 100 * agree.mean(), ulp, 100 * same_rows.mean()))
        assert ulp <= 2
        if not fma:
            assert same_rows[agree].all(), "decode differs although both exponentials agree"
        # everywhere: the score / -1 pattern (min-size and valid-range rules) and the boxes to 1e-3 px
        assert np.abs(boxes[:, :4] - res["dets"][:, :4]).max() < 2e-3
        flip = (boxes[:, 4] == -1) != (res["dets"][:, 4] == -1)
        assert flip.mean() < 1e-4 and not flip[agree & same_rows].any()
        if fma:
            assert same_rows.mean() > 0.5        # FMA contraction changes last bits of some rows, never the rules


@pytest.mark.parametrize("seed,B,H,tie,dead", [(1, 3, 32, 0.0, False), (3, 20, 32, 0.0, False), (4, 4, 32, 0.5, False),
                                             (5, 2, 8, 0.3, True)])
def test_nms_kernel_tie_order_fillers_and_gt_assignment(seed, B, H, tie, dead):
    """Same decoded rows in -> the reference NMS kernel, the oracle and the product op must emit the same bytes."""
    from sniper_b200 import ops
    inp = CASES.mpt_case(seed, B, H, H, tie, dead)
    cls_prob, bbox_pred, im_info, gts, vr = inp
    res, rois_pre, nk = _oracle_stages(inp, H, H)
    for fma in (False, True):      # no floating-point contraction can occur in the IoU test's compare chain outcome
        out = R.nms(_t(res["dets"]), B, A, H, H, 300, fma).cpu().numpy()
        assert out.tobytes() == rois_pre.tobytes(), "reference NMS kernel and oracle disagree (fma=%d)" % fma
    if dead:
        assert nk[-1] == 0
    ref = R.host_assign(gts, out, vr)                       # the reference's own host half on the reference NMS output
    assert ref["rois"].tobytes() == res["rois"].tobytes() and np.array_equal(ref["label"], res["label"])
    assert ref["bbox_target"].tobytes() == res["bbox_target"].tobytes()
    # product kernel (decode bit-identical to the oracle's): whole operator == reference kernels + reference host code
    t = [_t(a) for a in inp]
    rois, label, bt, bw, keep, nkept = ops.multi_proposal_target(*t, return_keep=True)
    assert rois.cpu().numpy().tobytes() == ref["rois"].tobytes()
    assert np.array_equal(label.cpu().numpy(), ref["label"]) and np.array_equal(bw.cpu().numpy(), ref["bbox_weight"])
    assert np.array_equal(nkept.cpu().numpy(), nk)
    d = np.abs(bt.cpu().numpy().view(np.int32).astype(np.int64) - ref["bbox_target"].view(np.int32).astype(np.int64))
    assert d.max() <= 1                                      # device double log vs libm: <= 1 ulp


def test_whole_reference_gpu_operator_vs_product():
    """getProps -> NMS -> host assignment, all reference code, against the product operator on the same inputs.  The two
    differ only through CUDA's expf in the decode (<= 2 ulp on some anchors)."""
    from sniper_b200 import ops
    B, H = 6, 32
    inp = CASES.mpt_case(11, B, H, H)
    cls_prob, bbox_pred, im_info, gts, vr = inp
    anchors = R.generate_anchors(STRIDE, CASES.SCALES, CASES.RATIOS)
    boxes = R.get_props(_t(bbox_pred), _t(im_info), _t(anchors), _t(cls_prob), _t(vr), B, A, H, H, STRIDE)
    props = R.nms(boxes, B, A, H, H).cpu().numpy()
    ref = R.host_assign(gts, props, vr)
    rois, label, bt, bw = ops.multi_proposal_target(*[_t(a) for a in inp])
    r = rois.cpu().numpy()
    same = (r.view(np.int32) == ref["rois"].view(np.int32)).all(1)
    near = np.abs(r - ref["rois"]).max(1) < 2e-3
    lab_same = (label.cpu().numpy() == ref["label"]).mean()
    print("roi rows: %d identical, %d within 2e-3 px, of %d; labels equal %.4f" % (same.sum(), near.sum(), same.size, lab_same))
    # a 1-ulp difference in a kept box can flip a later `ovr > 0.7` decision and shift every following row of that chip,
    # so the bound is on the fraction of rows, per chip the prefix before the first divergence must be identical
    assert near.mean() > 0.9 and lab_same > 0.97
    for b in range(B):
        blk = near[b * 300:(b + 1) * 300]
        first_bad = int(np.argmin(blk)) if not blk.all() else 300
        assert first_bad > 20, (b, first_bad)


def test_deformable_psroi_kernels():
    from sniper_b200 import ops
    for name, data, rois, trans, kw in CASES.dpsroi_cases():
        for no_trans in (True, False):
            tr = None if no_trans else trans
            out, cnt, _ = O.deform_psroi_fwd(data, rois, tr, no_trans=no_trans, **kw)
            r_out, r_cnt = R.dpsroi_fwd(_t(data), _t(rois), None if no_trans else _t(trans), **kw)
            assert np.array_equal(r_cnt.cpu().numpy(), cnt), name
            assert r_out.cpu().numpy().tobytes() == out.tobytes(), name       # reference kernel == oracle, bit for bit
            f_out, f_cnt = R.dpsroi_fwd(_t(data), _t(rois), None if no_trans else _t(trans), fma=True, **kw)
            assert np.array_equal(f_cnt.cpu().numpy(), cnt)
            dfma = np.abs(f_out.cpu().numpy() - out).max()
            print("%s no_trans=%d: FMA build max |delta| %.3e" % (name, no_trans, dfma))
            assert dfma <= 4e-6 * np.abs(data).max()
            g_out, g_cnt, _ = ops.deform_psroi_fwd(_t(data), _t(rois), None if no_trans else _t(trans),
                                                   spatial_scale=kw["spatial_scale"], output_dim=kw["output_dim"],
                                                   group_size=kw["group_size"], pooled_size=kw["pooled"],
                                                   part_size=kw["part_size"], sample_per_part=kw["spp"],
                                                   trans_std=kw["trans_std"], no_trans=no_trans, want_sample_idx=True)
            assert g_out.cpu().numpy().tobytes() == r_out.cpu().numpy().tobytes()   # product (generic kernel) == reference
            # backward: float atomics -> compare with the oracle's double accumulation
            g = np.random.RandomState(1).randn(*out.shape).astype(np.float32)
            dd, td = O.deform_psroi_bwd(g, cnt, data, rois, tr, no_trans=no_trans, **kw)
            r_dd, r_td = R.dpsroi_bwd(_t(g), r_cnt, _t(data), _t(rois), None if no_trans else _t(trans), **kw)
            np.testing.assert_allclose(r_dd.cpu().numpy(), dd, rtol=1e-4, atol=2e-5 * np.abs(g).max() * 16)
            if not no_trans:
                np.testing.assert_allclose(r_td.cpu().numpy(), td, rtol=1e-3, atol=1e-4 * np.abs(td).max())


def test_psroi_kernels():
    from sniper_b200 import ops
    data, rois, kw = CASES.psroi_case()
    out, bins = O.psroi_fwd(data, rois, **kw)
    r_out = R.psroi_fwd(_t(data), _t(rois), **kw)
    assert r_out.cpu().numpy().tobytes() == out.tobytes()
    g_out, _ = ops.psroi_fwd(_t(data), _t(rois), spatial_scale=kw["spatial_scale"], output_dim=kw["output_dim"],
                             group_size=kw["group_size"], pooled_size=kw["pooled"])
    assert g_out.cpu().numpy().tobytes() == out.tobytes()
    g = np.random.RandomState(2).randn(*out.shape).astype(np.float32)
    dd = O.psroi_bwd(g, rois, data.shape, **kw)
    r_dd = R.psroi_bwd(_t(g), _t(rois), data.shape, **kw)
    np.testing.assert_allclose(r_dd.cpu().numpy(), dd, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("big", [True, False])
def test_deformable_im2col_kernels(big):
    """Reference im2col / col2im / col2im_coord vs the product's NHWC kernels and the float64 restatement used by the
    graph-parity test (oracle/torch_graph.deform_conv2d).  Layouts: reference col [N, C*9, H, W] (channel-major (c,tap)),
    product col [N*H*W, 9*C] (tap-major)."""
    import torch
    from sniper_b200 import ops
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch_graph as TG
    x, off, dcol = CASES.deform_case(big_offsets=big)
    N, C, H, _ = x.shape
    r_col = R.deform_im2col(_t(x), _t(off))                                       # [N, C*9, H, W]
    xh = _t(x.transpose(0, 2, 3, 1))
    offh = torch.zeros(N, H, H, 96, device="cuda")
    offh[..., :72] = _t(off.transpose(0, 2, 3, 1))
    p_col = ops.deform_im2col(xh, offh, kh=3, kw=3, stride=1, dil=2, pad=2, dgroups=4)   # [N*H*W, 9*C]
    r_as_p = r_col.view(N, C, 9, H, H).permute(0, 3, 4, 2, 1).reshape(N * H * H, 9 * C)
    err = (p_col - r_as_p).abs().max().item()
    print("im2col product vs reference max |delta| %.3e" % err)
    assert err <= 2e-6 * np.abs(x).max()
    nz = (r_as_p != 0)
    assert torch.equal(nz, p_col != 0) or ((nz != (p_col != 0)).float().mean().item() < 1e-5)   # same in/out-of-image rule
    # float64 restatement: identity weight turns deform_conv2d into im2col of one tap-channel at a time
    xd, od = _t(x).double(), _t(off).double()
    w = torch.zeros(9 * C, C, 3, 3, dtype=torch.float64, device="cuda")
    for tap in range(9):
        w[tap * C + torch.arange(C), torch.arange(C), tap // 3, tap % 3] = 1.0
    col64 = TG.deform_conv2d(xd, od, w)                                           # [N, 9*C, H, W] tap-major
    c64 = col64.view(N, 9, C, H, H).permute(0, 3, 4, 1, 2).reshape(N * H * H, 9 * C)
    assert (c64 - r_as_p.double()).abs().max().item() <= 2e-6 * np.abs(x).max()
    # gradients
    r_gi, r_go = R.deform_col2im(_t(dcol), _t(x), _t(off))
    dcol_p = _t(dcol).view(N, C, 9, H, H).permute(0, 3, 4, 2, 1).reshape(N * H * H, 9 * C).contiguous()
    p_gi, p_go = ops.deform_col2im(dcol_p, xh, offh, kh=3, kw=3, stride=1, dil=2, pad=2, dgroups=4)
    scale = float(np.abs(dcol).max() * 9)
    assert (p_gi.permute(0, 3, 1, 2) - r_gi).abs().max().item() <= 1e-5 * scale
    e_off = (p_go[..., :72].permute(0, 3, 1, 2) - r_go).abs().max().item()
    print("col2im_coord product vs reference max |delta| %.3e (|grad| max %.3e)" % (e_off, r_go.abs().max().item()))
    assert e_off <= 2e-5 * r_go.abs().max().item() + 1e-5


def test_deformable_psroi_full_config_vs_reference_kernel():
    """The two pooling calls of the ResNet-101 head at the metric's size (B = 20 chips, 6000 rois, 256 channels, 7x7 bins,
    4x4 samples; resnet_mx_101_e2e.py:286-293) on the product's NHWC hot-path kernels against the reference's own CUDA
    kernels run on the same GPU: sample counts exact, oracle-order forward (SNIPER_PSROI_EXACT=1) bit-exact, default
    separable forward <= 2e-6 * max|x|, backward (float atomics on both sides) to 2e-4 of the largest gradient."""
    import torch
    from sniper_b200 import ops, synth
    rng = np.random.RandomState(17)
    B, C, N = 20, 256, 6000
    data = rng.randn(B, C, 32, 32).astype(np.float32)
    rois = synth.rois_for_pool(rng, N, B)
    rois[:, 0] = np.repeat(np.arange(B), N // B)                      # 300 rois per chip, as MultiProposalTarget emits them
    trans = (rng.randn(N, 2, 7, 7) * 0.3).astype(np.float32)
    kw = dict(spatial_scale=0.0625, output_dim=256, group_size=1, pooled=7, part_size=7, spp=4, trans_std=0.1)
    pkw = dict(spatial_scale=0.0625, output_dim=256, group_size=1, pooled_size=7, part_size=7, sample_per_part=4,
               trans_std=0.1, layout=ops.NHWC)
    d_nchw, d_nhwc = _t(data), _t(data.transpose(0, 2, 3, 1))
    r, tr = _t(rois), _t(trans)
    g = _t(rng.randn(N, 256, 7, 7).astype(np.float32))
    g_nhwc = g.permute(0, 2, 3, 1).contiguous()
    for no_trans in (True, False):
        t_ = None if no_trans else tr
        r_out, r_cnt = R.dpsroi_fwd(d_nchw, r, t_, **kw)
        os.environ["SNIPER_PSROI_EXACT"] = "1"
        try:
            e_out, e_cnt, _ = ops.deform_psroi_fwd(d_nhwc, r, t_, no_trans=no_trans, **pkw)
        finally:
            os.environ.pop("SNIPER_PSROI_EXACT", None)
        assert torch.equal(e_cnt.permute(0, 3, 1, 2), r_cnt)
        assert torch.equal(e_out.permute(0, 3, 1, 2), r_out), "oracle-order NHWC kernel differs from the reference kernel"
        s_out, _, _ = ops.deform_psroi_fwd(d_nhwc, r, t_, no_trans=no_trans, want_count=False, **pkw)
        err = (s_out.permute(0, 3, 1, 2) - r_out).abs().max().item()
        assert err <= 2e-6 * np.abs(data).max(), err
        r_dd, r_td = R.dpsroi_bwd(g, r_cnt, d_nchw, r, t_, **kw)
        p_dd, p_td = ops.deform_psroi_bwd(g_nhwc, d_nhwc, r, t_, no_trans=no_trans, **pkw)
        e_dd = (p_dd.permute(0, 3, 1, 2) - r_dd).abs().max().item() / r_dd.abs().max().item()
        print("no_trans=%d: separable fwd max err %.2e, bwd data rel err %.2e" % (no_trans, err, e_dd))
        assert e_dd < 2e-4
        if not no_trans:
            e_td = (p_td - r_td).abs().max().item() / r_td.abs().max().item()
            print("           bwd trans rel err %.2e" % e_td)
            assert e_td < 1e-3
