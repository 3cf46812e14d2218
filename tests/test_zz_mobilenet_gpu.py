"""MobileNetV2 SNIPER path (BASELINE config 4) on the GPU, through the C-ABI:
  * depthwise 3x3 forward / data gradient / weight gradient, first-layer im2col, shortcut add against float64 torch
    (the same comparisons tests/test_depthwise_cpu.py makes for the per-thread bodies executed on the CPU);
  * BatchNorm backward with the clip(0, 6) / no-activation masks against autograd; affine_act's clip;
  * one whole training step against the float64 restatement of the reference's mobilenetv2_e2e graph
    (oracle/torch_graph_mnv2.py, parameter set == the graph the reference's own symbol file builds): rois / labels
    bit-exact vs the C oracle on the step's own RPN outputs, activations, the four loss sums, EVERY parameter gradient;
  * the trainer (CUDA graphs, schedule, update) on this network.
Tolerances: fp32 kernels 1e-5 relative; TF32 / bf16 graph tolerances as in tests/test_graph_parity_gpu.py (stated at the
asserts; the measured figures are printed)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("NB,H,W,C,stride", [(2, 64, 64, 64, 1), (2, 64, 64, 128, 2), (3, 17, 23, 192, 1), (1, 15, 9, 8, 2),
                                               (4, 256, 256, 64, 1), (2, 16, 16, 960, 1)])
def test_depthwise_kernels_match_float64(dtype, NB, H, W, C, stride):
    import torch
    import torch.nn.functional as F
    from sniper_b200 import ops
    dt = torch.float32 if dtype == "fp32" else torch.bfloat16
    torch.manual_seed(H + C + stride)
    x = torch.randn(NB, H, W, C, device="cuda").to(dt)
    w = torch.randn(9, C, device="cuda") * 0.3
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    dy = torch.randn(NB, Ho, Wo, C, device="cuda").to(dt)
    xd = x.double().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    wd = w.double().t().reshape(C, 1, 3, 3).contiguous().requires_grad_(True)
    yr = F.conv2d(xd, wd, stride=stride, padding=1, groups=C)
    yr.backward(dy.double().permute(0, 3, 1, 2))
    tol = 1e-5 if dtype == "fp32" else 1e-2        # bf16: rounding of the stored output only (inputs are exact)
    y = ops.depthwise3x3(x, w, stride)
    assert torch.allclose(y.double(), yr.detach().permute(0, 2, 3, 1), rtol=tol, atol=3 * tol)
    dx = ops.depthwise3x3_dgrad(dy, w, (H, W), stride)
    assert torch.allclose(dx.double(), xd.grad.permute(0, 2, 3, 1), rtol=tol, atol=3 * tol)
    dw = torch.ones(9, C, device="cuda")
    ops.depthwise3x3_wgrad(x, dy, dw, stride)
    ref = wd.grad.reshape(C, 9).t()
    assert (dw.double() - 1.0 - ref).abs().max().item() <= 2e-4 * (ref.abs().max().item() + 1.0)   # fp32 atomics
    torch.cuda.synchronize()


def test_first_layer_im2col_and_add_rows():
    import torch
    import torch.nn.functional as F
    from sniper_b200 import ops
    torch.manual_seed(2)
    x = torch.randn(2, 3, 64, 96, device="cuda") * 50
    for dt in (torch.float32, torch.bfloat16):
        col = ops.im2col3x3s2(x, 64, dtype=dt)
        u = F.unfold(x, 3, padding=1, stride=2).view(2, 3, 9, 32 * 48).permute(0, 3, 2, 1).reshape(2, 32, 48, 27)
        assert torch.equal(col[..., :27].float(), u.to(dt).float()) and not col[..., 27:].float().any()
        a = torch.randn(1000, 192, device="cuda").to(dt)
        b = torch.randn(1000, 192, device="cuda").to(dt)
        want = (a.float() + b.float()).to(dt)
        assert torch.equal(ops.add_rows(a, b).float(), want.float())
        ops.add_rows(a, b, out=a)
        assert torch.equal(a.float(), want.float())
    # the GEMM on the im2col rows == the convolution (TF32)
    wt = torch.randn(64, 3, 3, 3, device="cuda") * 0.2
    rows = torch.zeros(64, 64, device="cuda")
    rows[:, :27] = wt.permute(0, 2, 3, 1).reshape(64, 27)
    col = ops.im2col3x3s2(x, 64)
    y = ops.conv2d_nhwc(col, rows, kh=1, kw=1)
    ref = F.conv2d(x.double(), wt.double(), stride=2, padding=1).permute(0, 2, 3, 1)
    assert _rel(y, ref) < 2e-3


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("act", [1, 2, 3])
def test_bn_backward_activation_variants(dtype, act):
    import torch
    from sniper_b200 import ops
    dt = torch.float32 if dtype == "fp32" else torch.bfloat16
    torch.manual_seed(10 + act)
    M, C = 5000, 192
    x = (torch.randn(M, C, device="cuda") * 2 + 0.5).to(dt)
    dy = torch.randn(M, C, device="cuda").to(dt)
    add = torch.randn(M, C, device="cuda").to(dt)
    st = ops.BNState(C, "cuda")
    st.gamma.uniform_(1.5, 3.0)          # wide outputs so that the clip at 6 is active
    st.beta.normal_(2.0, 1.0)
    st.dgamma, st.dbeta = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    ops.bn_stats(x, st, eps=1e-5, momentum=0.9)
    y = ops.affine_act(x, st.scale, st.shift, relu={1: 1, 2: 2, 3: 0}[act])
    xd = x.double().requires_grad_(True)
    yr = torch.nn.functional.batch_norm(xd, None, None, st.gamma.double(), st.beta.double(), True, 0.0, 1e-5)
    yr = {1: torch.relu, 2: lambda t: t.clamp(0, 6), 3: lambda t: t}[act](yr)
    yr.backward(dy.double())
    tol = 2e-5 if dtype == "fp32" else 2e-2
    assert _rel(y, yr) < tol
    if act == 2:
        assert float((y.float() == 6).float().mean()) > 0.01 and float((y.float() == 0).float().mean()) > 0.001
    dx = ops.bn_act_bwd(x, dy, st, act, add=add)
    # elements within rounding of a clip boundary may take either side: compare in norm
    assert _rel(dx, xd.grad + add.double()) < (2e-3 if dtype == "fp32" else 3e-2)
    assert not st.sums.any()             # left zeroed for the next use
    assert st.dbeta.abs().sum().item() > 0 and st.dgamma.abs().sum().item() > 0


def _build(B, bf16, seed=5):
    import torch
    from sniper_b200 import model_mnv2 as MM
    from sniper_b200 import synth_batch
    cfg = MM.MCfg()
    cfg.batch_images = B
    cfg.bf16 = bool(bf16)
    net = MM.SniperMobileNetV2(cfg, device="cuda", seed=seed)
    g = torch.Generator(device="cuda")
    g.manual_seed(seed + 1)
    for bn in net.all_bns():
        bn.st.gamma[:bn.C] = torch.empty(bn.C, device="cuda").uniform_(0.8, 1.2, generator=g)
        bn.st.beta[:bn.C] = torch.empty(bn.C, device="cuda").normal_(0, 0.1, generator=g)
    batch = synth_batch.make_batch(B, seed=7, device="cuda", A=cfg.num_anchors, stride=cfg.feat_stride)
    return cfg, net, batch


# Whole-graph tolerances.  This network is in the chaotic regime at random initialisation: BatchNorm re-normalises every
# one of its 53 layers, so a relative perturbation is carried through the whole depth and grows ~170x from the first
# layer to last_fm (measured on the float64 oracle: 2.4e-6 noise per 1x1 conv -> 4.2e-4, 1e-4 -> 1.7e-2, TF32 operand
# truncation itself -> 5.2e-2 against exact arithmetic).  TF32 / bf16 quantisation is discontinuous, so two evaluations
# whose inputs differ in the last fp32 bit re-quantise differently and settle at a fraction of the quantisation effect:
# measured on B200 (profiles/config4_r02.md) TF32: last_fm 1.4e-2 (3.7x closer to the TF32 model than that model is to
# exact arithmetic), loss sums 3e-5, gradients 0.24 median / 0.27 worst with norm ratios within 8 %; bf16: last_fm 0.20,
# loss sums 2e-3, gradients decorrelated (0.85) -- elementwise gradient parity in bf16 is therefore asserted per UNIT
# (test_inverted_residual_units_match_float64, teacher-forced inputs), and exactly on the CPU for the orchestration
# (tests/test_mnv2_wiring_cpu.py: float64 1e-7; bf16 and TF32 emulations bit-exact against the oracle's modes, and a
# float32-level seed in those exact evaluations reproduces the figures above: 1.36e-2 / 0.24 in TF32, 0.21 / 0.88 in bf16).  A wiring error shows as >= 0.7 with norm ratios far from 1.
TOL = {False: dict(act=4e-2, loss=1e-2, grad=0.45, median=0.35, head=0.25, norm=0.15),
       True: dict(act=0.45, loss=3e-2, grad=None, median=None, head=None, norm=None)}


def _unit_oracle(TM, TG, P, x64, prefix, cin, e, stride, shortcut, eps=1e-5):
    a1 = TM._unit(P, x64, prefix + "-exp", 1, 1, 1, True, eps)
    a2 = TM._unit(P, a1, prefix + "-depthwise", 3, stride, e, True, eps)
    y = TM._unit(P, a2, prefix + "-linear", 1, 1, 1, False, eps)
    return TG.qs(y + x64) if shortcut else y


@pytest.mark.parametrize("bf16", [False, True])
def test_inverted_residual_units_match_float64(bf16):
    """Each kind of inverted residual unit alone (forward, data gradient, the three weight gradients) on the product's
    kernels against the oracle's unit in the matching arithmetic mode, from the SAME input and output gradient: t = 1
    without shortcut, stride 2, shortcut, and the widest (960 -> 320).  No depth, so no amplification: tolerances are
    the rounding of one unit (TF32: 3 truncated contractions; bf16: 6 stored tensors)."""
    import torch
    import torch_graph as TG
    import torch_graph_mnv2 as TM
    from sniper_b200 import ops
    cfg, net, _ = _build(2, bf16)
    dt = torch.bfloat16 if bf16 else torch.float32
    arg, aux = net.export_reference()
    P, _ = TM.params_to_torch(arg, aux, torch.float64, "cuda")
    tol = dict(y=5e-3, dx=2e-2, gw=2e-2) if bf16 else dict(y=2e-4, dx=2e-3, gw=3e-3)      # measured: bf16 5e-4 / 2e-3 / 1.5e-3, TF32 1e-5 / 1e-4 / 2e-4
    torch.manual_seed(11)
    for idx in (0, 1, 2, 16):
        u = net.units[idx]
        cinp, coutp = u.exp.cin, u.lin.coutp
        H = 32
        x = torch.zeros(2, H, H, cinp, device="cuda")
        x[..., :u.cin] = (torch.randn(2, H, H, u.cin, device="cuda") + 1.0).clamp(0, 6)
        x = x.to(dt)
        ops.weight_transpose_batched(ops.weight_transpose_jobs([j for c in u.convs() for j in c.bwd_jobs()], "cuda"))
        net.P.g.zero_()
        y = u.fwd(x, cfg)
        Ho = y.shape[1]
        dy = torch.zeros(2, Ho, Ho, coutp, device="cuda")
        dy[..., :u.cout] = torch.randn(2, Ho, Ho, u.cout, device="cuda")
        dy = dy.to(dt)
        dx = u.bwd(dy, cfg)
        ops.bn_param_grad_batched(ops.bn_param_grad_jobs([b.st for b in u.bns()], "cuda"))
        cfg.wsched.join()
        torch.cuda.synchronize()
        assert not y[..., u.cout:].float().any() and not dx[..., u.cin:].float().any()          # padding stays zero
        x64 = x[..., :u.cin].double().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
        for p in P.values():
            p.grad = None
        TG.MODE[0] = "bf16" if bf16 else "tf32"
        TG.LOWP[0] = True
        try:
            yr = _unit_oracle(TM, TG, P, x64, u.prefix, u.cin, u.e, u.stride, u.shortcut)
            yr.backward(dy[..., :u.cout].double().permute(0, 3, 1, 2))
        finally:
            TG.MODE[0] = "exact"
            TG.LOWP[0] = False
        garg, _ = net.export_reference(grads=True)
        errs = dict(y=_rel(y[..., :u.cout].permute(0, 3, 1, 2), yr), dx=_rel(dx[..., :u.cin].permute(0, 3, 1, 2), x64.grad))
        for part in ("-exp", "-depthwise", "-linear"):
            name = u.prefix + part + "-conv2d_weight"
            errs["gw" + part] = _rel(torch.from_numpy(garg[name]).cuda(), P[name].grad)
        print(u.prefix, "bf16" if bf16 else "tf32", {k: "%.2e" % v for k, v in errs.items()})
        for k, v in errs.items():
            assert v < tol[k[:2]], (u.prefix, k, v)


@pytest.mark.parametrize("bf16", [False, True])
def test_training_graph_matches_float64_reference(bf16):
    import torch
    import oracle_lib as O
    import torch_graph as TG
    import torch_graph_mnv2 as TM
    B = 2
    cfg, net, batch = _build(B, bf16)
    out = net.forward_backward(batch)
    torch.cuda.synchronize()
    A = cfg.num_anchors
    prob = out["rpn_cls_prob"].permute(0, 3, 1, 2).contiguous()
    bbox = out["rpn_head"][..., :4 * A].permute(0, 3, 1, 2).contiguous()
    res = O.multi_proposal_target(prob.cpu().numpy(), bbox.cpu().numpy(), batch["im_info"].cpu().numpy(),
                                  batch["gt_boxes"].cpu().numpy(), batch["valid_ranges"].cpu().numpy(), feat_stride=32,
                                  scales=cfg.scales, ratios=cfg.ratios)
    assert out["rois"].cpu().numpy().tobytes() == res["rois"].tobytes()
    assert np.array_equal(out["label"].cpu().numpy(), res["label"].reshape(-1))
    assert int((res["label"] > 0).sum()) > 0, "test batch yields no foreground roi"
    # padded channels stay exactly zero
    assert not out["first"][..., cfg.first_c:].float().any()
    for c in net.backbone_convs():
        gw = net.P.grad(c.name + "_weight")
        assert not gw[c.cout:].any() and not gw[:, c.cin_real:].any(), c.name
    arg, aux = net.export_reference()
    garg, _ = net.export_reference(grads=True)
    P, Aux = TM.params_to_torch(arg, aux, torch.float64, "cuda")
    b64 = {k: v.double() for k, v in batch.items()}
    TG.MODE[0] = "bf16" if bf16 else "tf32"
    try:
        obj, ref = TM.forward_train(P, Aux, b64, lambda *_: res, batch_images=B)
        obj.backward()
    finally:
        TG.MODE[0] = "exact"
        TG.LOWP[0] = False
    tol = TOL[bool(bf16)]
    acts = dict(last_fm=_rel(out["last_fm"].permute(0, 3, 1, 2), ref["last_fm"]), rpn_prob=_rel(prob, ref["rpn_cls_prob"]),
                rpn_bbox=_rel(bbox, ref["rpn_bbox_pred"]), cls_prob=_rel(out["cls_prob"], ref["cls_prob"]))
    print("activation errors (%s):" % ("bf16" if bf16 else "tf32"), {k: "%.2e" % v for k, v in acts.items()})
    ls, lr = out["losses"][:4].double().cpu(), ref["loss_sums"].cpu()
    print("losses ours", ls.tolist(), "reference", lr.tolist())
    rows = []
    for name, p in P.items():
        if not p.requires_grad:
            assert name not in garg
            continue
        ours = torch.from_numpy(garg[name]).cuda()
        assert ours.shape == p.grad.shape, (name, ours.shape, p.grad.shape)
        rows.append((_rel(ours, p.grad), name, p.grad.norm().item(), ours.double().norm().item()))
    rows.sort(reverse=True)
    med = rows[len(rows) // 2][0]
    print("worst gradient errors:", [(round(r[0], 4), r[1]) for r in rows[:8]], "median %.2e over %d tensors" % (med, len(rows)))
    assert len(rows) == 71
    for k, v in acts.items():
        assert v < tol["act"], (k, v)
    for i in range(4):
        assert abs(ls[i] - lr[i]) <= tol["loss"] * abs(lr[i]) + 1e-4, (i, ls[i].item(), lr[i].item())
    ratios = [ours_n / nrm for _, _, nrm, ours_n in rows if nrm > 1e-9]
    print("gradient norm ratios: min %.3f max %.3f" % (min(ratios), max(ratios)))
    assert 0.5 < min(ratios) and max(ratios) < 2.0
    if tol["grad"] is not None:
        for r, name, nrm, ours_n in rows:
            assert r < tol["grad"] or nrm < 1e-9, (name, r, nrm)
            assert abs(ours_n / nrm - 1) < tol["norm"] or nrm < 1e-9, (name, ours_n, nrm)
            if "seq-" not in name and "first" not in name and "last" not in name:
                assert r < tol["head"], (name, r)
        assert med < tol["median"]


def test_trainer_runs_the_mobilenet_graph_under_cuda_graphs():
    import torch
    from sniper_b200 import model_mnv2 as MM
    from sniper_b200 import synth_batch
    from sniper_b200.trainer import Trainer
    cfg = MM.MCfg()
    cfg.batch_images = 2
    cfg.bf16 = True
    cfg.lr = 0.01
    net = MM.SniperMobileNetV2(cfg, device="cuda:0", seed=3)
    tr = Trainer(cfg, device="cuda:0", use_graph=True, net=net, scheduler=None)       # constant lr
    host = [synth_batch.make_batch(2, seed=20 + i, device="cpu", pinned=True, A=cfg.num_anchors, stride=cfg.feat_stride)
            for i in range(2)]
    w0 = net.P.w.clone()
    losses = [tr.step(host[i % 2], prefetch=host[(i + 1) % 2]) for i in range(6)]
    torch.cuda.synchronize()
    vals = np.array([[l["rpn_cls_loss"], l["rpn_bbox_loss"], l["rcnn_cls_loss"], l["rcnn_bbox_loss"]] for l in losses])
    print("losses per step", vals.tolist(), "lr", [l["lr"] for l in losses])
    assert np.isfinite(vals).all() and tr.g_fb is not None and len(tr.g_fb) == 1
    assert not torch.equal(net.P.w, w0) and torch.isfinite(net.P.w).all()
    assert vals[-2:, 2].mean() < vals[:2, 2].mean()          # the R-CNN classification loss goes down on two repeating batches
    for bn in net.all_bns():                                  # fixed parameters did not move; padding stayed zero
        o, _ = net.P.layout[bn.name + "_gamma"]
        assert torch.equal(net.P.w[o:o + bn.Cp], w0[o:o + bn.Cp])
    for c in net.backbone_convs():
        w = net.P[c.name + "_weight"]
        assert not w[c.cout:].any() and not w[:, c.cin_real:].any()


def test_inference_forward_path():
    """mobilenetv2_e2e.get_symbol_rcnn(is_train=False) as SniperMobileNetV2.forward_inference: proposals from the device
    MultiProposal op at stride 32, class probabilities that sum to one, boxes inside the chip, bit-identical on a second
    call, parameters untouched (the comparison with the float64 test graph runs on the CPU: test_mnv2_wiring_cpu.py)."""
    import torch
    cfg, net, batch = _build(2, True)
    net.train_step(batch, lr=0.001)
    w0 = net.P.w.clone()
    rois, scores, cls_prob, bbox_pred = net.forward_inference(batch["data"], batch["im_info"])
    torch.cuda.synchronize()
    assert rois.shape == (600, 5) and scores.shape == (600,) and cls_prob.shape == (600, 81) and bbox_pred.shape == (600, 4)
    assert torch.isfinite(cls_prob).all() and torch.isfinite(bbox_pred).all()
    assert (cls_prob.sum(1) - 1).abs().max().item() < 1e-4
    assert (rois[:, 1:] >= 0).all() and (rois[:, 1:] <= 511).all()
    r2 = net.forward_inference(batch["data"], batch["im_info"])
    assert all(torch.equal(a, b) for a, b in zip((rois, scores, cls_prob, bbox_pred), r2))
    assert torch.equal(w0, net.P.w)
