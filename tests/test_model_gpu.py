"""Layer-composition parity on the GPU: data gradients, residual units (train-mode BN), deformable conv and
a whole SNIPER training step, against PyTorch fp64 autograd references of the same graph.

Tolerances: TF32 products (10-bit mantissa) -> relative 2^-10 per product; asserted as relative Frobenius
error <= 3e-3 for activations/gradients of single layers; after a full residual unit (6 chained TF32
contractions, three BN+ReLU whose masks can flip on near-zero pre-activations) <= 4e-2 Frobenius with the
median element-wise relative error <= 3e-3.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def _mk_conv(cin, cout, k, stride, dil, pad, bias=False, cout_pad=None, plain_wt=False):
    import torch
    from sniper_b200 import model
    P = model.ParamStore()
    c = model.Conv(P, "c", cin, cout, k, stride, dil, pad, bias=bias, cout_pad=cout_pad, plain_wt=plain_wt)
    P.finalize("cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    c.init(device="cuda", gen=g)
    return P, c


@pytest.mark.parametrize("cin,cout,k,stride,dil,pad,H", [(64, 128, 1, 1, 1, 0, 32), (128, 128, 3, 1, 1, 1, 32),
                                                       (128, 96, 3, 1, 2, 2, 32), (128, 128, 3, 2, 1, 1, 64),
                                                       (256, 512, 1, 2, 1, 0, 64), (128, 128, 3, 2, 1, 1, 128)])
def test_conv_data_and_weight_gradients(cin, cout, k, stride, dil, pad, H):
    import torch
    import torch.nn.functional as F
    P, c = _mk_conv(cin, cout, k, stride, dil, pad)
    torch.manual_seed(0)
    x = torch.randn(2, H, H, cin, device="cuda")
    y = c.fwd(x)
    dy = torch.randn_like(y)
    c.prepare_bwd()
    dx = c.bwd_data(dy, (H, H))
    c.bwd_weight(dy, x, 4)
    xd = x.permute(0, 3, 1, 2).double().requires_grad_(True)
    wd = c.w.view(cout, k, k, cin).permute(0, 3, 1, 2).double().requires_grad_(True)
    yr = F.conv2d(xd, wd, None, stride=stride, padding=pad, dilation=dil)
    yr.backward(dy.permute(0, 3, 1, 2).double())
    assert _rel(y, yr.permute(0, 2, 3, 1)) < 3e-3
    assert _rel(dx, xd.grad.permute(0, 2, 3, 1)) < 3e-3
    assert _rel(P.grad("c_weight").view(cout, k, k, cin), wd.grad.permute(0, 2, 3, 1)) < 3e-3
    # accumulate-into-residual path
    r = torch.randn(2, H, H, cin, device="cuda")
    r0 = r.clone()
    dx2 = c.bwd_data(dy, (H, H), out=r, residual=r)
    assert _rel(dx2, xd.grad.permute(0, 2, 3, 1) + r0.double()) < 3e-3


def _torch_unit(u, x, dout, cfg, frozen=False):
    """fp64 autograd reference of Unit.fwd (non-deform)."""
    import torch
    import torch.nn.functional as F
    params = {}

    def W(c):
        w = c.w.view(c.coutp, c.k, c.k, c.cin).permute(0, 3, 1, 2).double().clone().requires_grad_(True)
        params[c.name] = w
        return w

    def bn(b, t):
        g = b.st.gamma.double().clone().requires_grad_(True)
        be = b.st.beta.double().clone().requires_grad_(True)
        params[b.name] = (g, be)
        return torch.relu(F.batch_norm(t, None, None, g, be, True, 0.0, cfg.bn_eps))

    xd = x.permute(0, 3, 1, 2).double().requires_grad_(True)
    a1 = bn(u.bn1, xd)
    c1 = F.conv2d(a1, W(u.conv1))
    a2 = bn(u.bn2, c1)
    c2 = F.conv2d(a2, W(u.conv2), stride=u.conv2.stride, padding=u.conv2.pad, dilation=u.conv2.dil)
    a3 = bn(u.bn3, c2)
    c3 = F.conv2d(a3, W(u.conv3))
    sc = xd if u.dim_match else F.conv2d(a1, W(u.sc), stride=u.sc.stride)
    y = c3 + sc
    y.backward(dout.permute(0, 3, 1, 2).double())
    return y.permute(0, 2, 3, 1), xd.grad.permute(0, 2, 3, 1), params


@pytest.mark.parametrize("cin,cout,stride,dim_match,H", [(256, 256, 1, True, 32), (128, 256, 2, False, 64)])
def test_residual_unit_train_bn(cin, cout, stride, dim_match, H):
    import torch
    from sniper_b200 import model
    cfg = model.Cfg()
    cfg.wgrad_splits = 4
    cfg.wsched = model.WgradScheduler(False)
    P = model.ParamStore()
    u = model.Unit(P, "u", cin, cout, stride, dim_match, frozen=False)
    P.finalize("cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(2)
    for c in u.convs():
        c.init(device="cuda", gen=g)
    for b in u.bns():
        b.build("cuda")
        b.st.gamma.uniform_(0.5, 1.5)
        b.st.beta.normal_(0, 0.2)
    torch.manual_seed(3)
    x = torch.randn(2, H, H, cin, device="cuda")
    y = u.fwd(x, cfg)
    dout = torch.randn_like(y)
    for c in u.convs():
        c.prepare_bwd()
    dx = u.bwd(dout, cfg)
    yr, dxr, params = _torch_unit(u, x, dout, cfg)
    assert _rel(y, yr) < 5e-3
    assert _rel(dx, dxr) < 4e-2
    med = ((dx.double() - dxr).abs() / (dxr.abs() + 1e-3)).median().item()
    assert med < 8e-3
    for c in u.convs():
        gr = params[c.name].grad.permute(0, 2, 3, 1).reshape(c.coutp, -1)
        assert _rel(P.grad(c.name + "_weight"), gr) < 4e-2, c.name
    for b in u.bns():
        gg, gb = params[b.name]
        assert _rel(P.grad(b.name + "_gamma"), gg.grad) < 4e-2, b.name
        assert _rel(P.grad(b.name + "_beta"), gb.grad) < 4e-2, b.name


def _torch_deform_conv(x, offset, w, dil=2, pad=2, dg=4):
    """fp64 restatement of DeformableConvolution forward (deformable_im2col.cuh:78-113,216-263 + GEMM),
    differentiable w.r.t. x, offset, w.  x [N,H,W,C], offset [N,H,W,dg*18], w [Cout,3,3,C]."""
    import torch
    N, H, W, C = x.shape
    cpg = C // dg
    hh, ww = torch.meshgrid(torch.arange(H, device=x.device), torch.arange(W, device=x.device), indexing="ij")
    cols = []
    for i in range(3):
        for j in range(3):
            tap = i * 3 + j
            per_g = []
            for g in range(dg):
                oh = offset[..., g * 18 + 2 * tap]
                ow = offset[..., g * 18 + 2 * tap + 1]
                h_im = (hh - pad + i * dil).to(x.dtype) + oh
                w_im = (ww - pad + j * dil).to(x.dtype) + ow
                valid = (h_im >= 0) & (w_im >= 0) & (h_im < H) & (w_im < W)
                h_low = torch.floor(h_im).clamp(max=H - 1)
                w_low = torch.floor(w_im).clamp(max=W - 1)
                hc = torch.where(torch.floor(h_im) >= H - 1, h_low, h_im)
                wc = torch.where(torch.floor(w_im) >= W - 1, w_low, w_im)
                h_high = (h_low + 1).clamp(max=H - 1)
                w_high = (w_low + 1).clamp(max=W - 1)
                lh, lw = hc - h_low, wc - w_low
                xg = x[..., g * cpg:(g + 1) * cpg]
                n_idx = torch.arange(N, device=x.device).view(N, 1, 1).expand(N, H, W)

                def at(hi, wi):
                    return xg[n_idx, hi.long().clamp(0, H - 1), wi.long().clamp(0, W - 1)]
                v = ((1 - lh) * (1 - lw)).unsqueeze(-1) * at(h_low, w_low) + ((1 - lh) * lw).unsqueeze(-1) * at(h_low, w_high) \
                    + (lh * (1 - lw)).unsqueeze(-1) * at(h_high, w_low) + (lh * lw).unsqueeze(-1) * at(h_high, w_high)
                per_g.append(v * valid.unsqueeze(-1).to(x.dtype))
            cols.append(torch.cat(per_g, -1))
    col = torch.stack(cols, 3)                                   # [N,H,W,9,C]
    return torch.einsum("nhwtc,otc->nhwo", col, w.reshape(w.shape[0], 9, C))


def test_deformable_conv_fwd_bwd():
    import torch
    from sniper_b200 import ops
    torch.manual_seed(5)
    N, H, C, Cout = 2, 16, 512, 128
    x = torch.randn(N, H, H, C, device="cuda")
    off = torch.zeros(N, H, H, 96, device="cuda")
    off[..., :72] = torch.randn(N, H, H, 72, device="cuda") * 1.5
    w = torch.randn(Cout, 3, 3, C, device="cuda") / (9 * C) ** 0.5
    col = ops.deform_im2col(x, off, kh=3, kw=3, stride=1, dil=2, pad=2, dgroups=4)
    y = ops.gemm_nt(col, w.reshape(Cout, -1).contiguous()).view(N, H, H, Cout)
    xd = x.double().requires_grad_(True)
    od = off[..., :72].double().requires_grad_(True)
    wd = w.double().requires_grad_(True)
    yr = _torch_deform_conv(xd, od, wd)
    assert _rel(y, yr) < 3e-3
    dy = torch.randn_like(y)
    yr.backward(dy.double())
    M = N * H * H
    wt = w.reshape(Cout, -1).t().contiguous()                    # [9C, Cout]
    dcol = ops.gemm_nt(dy.view(M, Cout), wt)
    dx, doff = ops.deform_col2im(dcol, x, off, kh=3, kw=3, stride=1, dil=2, pad=2, dgroups=4)
    assert _rel(dx, xd.grad) < 3e-3
    assert _rel(doff[..., :72], od.grad) < 3e-3
    assert float(doff[..., 72:].abs().max()) == 0.0
    dw = ops.conv2d_wgrad_nhwc(dy, col.view(N, H, H, -1), kh=1, kw=1, splits=2)
    assert _rel(dw, wd.grad.reshape(Cout, -1)) < 3e-3


def test_full_training_step_smoke():
    """Two complete SNIPER training steps (B=2 chips of 512x512): finite losses, gradients reach every
    trainable tensor, frozen layers stay untouched, the loss decreases on a repeated batch."""
    import torch
    from sniper_b200 import model, synth_batch
    cfg = model.Cfg()
    cfg.batch_images = 2
    cfg.wgrad_splits = 4
    net = model.SniperResNet101(cfg, deform_offset_std=0.01)
    batch = synth_batch.make_batch(2, seed=7, device="cuda")
    w0 = net.P.w.clone()
    out = net.train_step(batch, lr=0.002)
    torch.cuda.synchronize()
    l0 = out["losses"].clone()
    assert torch.isfinite(l0).all() and torch.isfinite(net.P.g).all() and torch.isfinite(net.P.w).all()
    for name, gview in net.P.grads.items():
        if name.endswith("_bias") and "offset" in name:
            continue
        assert float(gview.abs().sum()) > 0, "no gradient reached " + name
    assert not torch.equal(w0, net.P.w)
    for _ in range(3):
        out = net.train_step(batch, lr=0.002)
    l1 = out["losses"].clone()
    assert torch.isfinite(l1).all()
    assert float(l1[2]) < float(l0[2])          # R-CNN classification loss goes down on the same batch
    assert out["rois"].shape == (600, 5) and out["cls_prob"].shape == (600, 81)


def test_fused_bn_statistics_match_separate_pass():
    """Column sums accumulated by the tcgen05 epilogue (stats=) == the colsum kernel's statistics."""
    import torch
    from sniper_b200 import model, ops
    torch.manual_seed(11)
    P, c = _mk_conv(128, 256, 3, 1, 1, 1)
    x = torch.randn(2, 32, 32, 128, device="cuda")
    a = ops.BNState(256, "cuda")
    b = ops.BNState(256, "cuda")
    y = c.fwd(x, stats=a.sums)
    ops.bn_finalize(a, y.numel() // 256, eps=2e-5, momentum=0.9)
    ops.bn_stats(y, b, eps=2e-5, momentum=0.9)
    assert (a.mean - b.mean).abs().max().item() < 1e-5
    assert ((a.invstd - b.invstd).abs() / b.invstd).max().item() < 1e-4
    assert float(a.sums.abs().sum()) == 0.0


def test_wgrad_side_stream_matches_single_stream():
    """Weight gradients scheduled on the second stream (WgradScheduler) == the single-stream backward, up to the
    run-to-run noise of the float atomics in the PSROI / col2im / split-K backward kernels (measured by running
    the single-stream configuration twice)."""
    import torch
    from sniper_b200 import model, synth_batch
    grads = []
    for side in (False, False, True):
        cfg = model.Cfg()
        cfg.batch_images = 1
        cfg.wgrad_splits = 1
        cfg.wgrad_stream = side
        net = model.SniperResNet101(cfg, deform_offset_std=0.01, seed=5)
        batch = synth_batch.make_batch(1, seed=9, device="cuda")
        net.forward_backward(batch)
        torch.cuda.synchronize()
        grads.append(net.P.g.clone())
    ref = float(grads[0].norm())
    noise = float((grads[0] - grads[1]).norm()) / ref
    diff = float((grads[0] - grads[2]).norm()) / ref
    print("run-to-run noise %.3e, side-stream difference %.3e" % (noise, diff))
    assert torch.isfinite(grads[2]).all()
    assert diff <= max(4.0 * noise, 1e-6), (noise, diff)


def test_trainer_prefetch_pipeline_delivers_the_right_batch():
    """Trainer.step(batch, prefetch=next): the H2D of the next batch runs on the copy stream into a staging set while the
    current step computes.  After every step the graph's static input buffers must hold exactly the batch that was
    passed to that step (prefetched or not), and the losses must be finite."""
    import math
    import torch
    from sniper_b200 import model, synth_batch, trainer
    batches = [synth_batch.make_batch(1, seed=s, device="cpu", pinned=True) for s in (1, 2, 3)]
    cfg = model.Cfg()
    cfg.batch_images = 1
    tr = trainer.Trainer(cfg, use_graph=True, seed=5)
    order = [0, 1, 2, 0, 2, 1, 1]
    for n, i in enumerate(order):
        nxt = batches[order[n + 1]] if n + 1 < len(order) and n != 3 else None      # step 4 is NOT prefetched
        losses = tr.step(batches[i], prefetch=nxt)
        assert all(math.isfinite(v) for v in losses.values())
        for k, v in batches[i].items():
            assert torch.equal(tr.static[k].cpu(), v), (n, k)


@pytest.mark.parametrize("cin,cout,stride,dim_match,H", [(256, 256, 1, True, 32), (128, 256, 2, False, 64)])
def test_residual_unit_inference_bn(cin, cout, stride, dim_match, H):
    """Unit.fwd_infer (BN with moving statistics folded into the conv epilogues) vs PyTorch fp64 eval-mode BN."""
    import torch
    import torch.nn.functional as F
    from sniper_b200 import model, ops
    cfg = model.Cfg()
    P = model.ParamStore()
    u = model.Unit(P, "u", cin, cout, stride, dim_match, frozen=False)
    P.finalize("cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(2)
    for c in u.convs():
        c.init(device="cuda", gen=g)
    for b in u.bns():
        b.build("cuda")
        b.st.gamma.uniform_(0.5, 1.5)
        b.st.beta.normal_(0, 0.2)
        b.st.moving_mean.normal_(0, 0.3)
        b.st.moving_var.uniform_(0.5, 2.0)
        ops.bn_frozen(b.st, cfg.bn_eps)
    torch.manual_seed(3)
    x = torch.randn(2, H, H, cin, device="cuda")
    y = u.fwd_infer(x, cfg)

    def W(c):
        return c.w.view(c.coutp, c.k, c.k, c.cin).permute(0, 3, 1, 2).double()

    def bn(b, t):
        return torch.relu(F.batch_norm(t, b.st.moving_mean.double(), b.st.moving_var.double(), b.st.gamma.double(),
                                       b.st.beta.double(), False, 0.0, cfg.bn_eps))
    xd = x.permute(0, 3, 1, 2).double()
    a1 = bn(u.bn1, xd)
    a2 = bn(u.bn2, F.conv2d(a1, W(u.conv1)))
    a3 = bn(u.bn3, F.conv2d(a2, W(u.conv2), stride=u.conv2.stride, padding=u.conv2.pad, dilation=u.conv2.dil))
    yr = F.conv2d(a3, W(u.conv3)) + (xd if dim_match else F.conv2d(a1, W(u.sc), stride=u.sc.stride))
    assert _rel(y, yr.permute(0, 2, 3, 1)) < 5e-3


def test_inference_forward_path():
    """get_symbol_rcnn(is_train=False) as SniperResNet101.forward_inference: proposals from the device MultiProposal op,
    class probabilities that sum to one, boxes inside the chip, bit-identical on a second call, parameters untouched."""
    import torch
    from sniper_b200 import model, synth_batch
    cfg = model.Cfg()
    cfg.batch_images = 2
    net = model.SniperResNet101(cfg, deform_offset_std=0.01)
    batch = synth_batch.make_batch(2, seed=7, device="cuda")
    net.train_step(batch, lr=0.001)                       # moving statistics become non-trivial
    w0 = net.P.w.clone()
    rois, scores, cls_prob, bbox_pred = net.forward_inference(batch["data"], batch["im_info"])
    torch.cuda.synchronize()
    assert rois.shape == (600, 5) and scores.shape == (600,) and cls_prob.shape == (600, 81) and bbox_pred.shape == (600, 4)
    assert torch.isfinite(cls_prob).all() and torch.isfinite(bbox_pred).all()
    assert (cls_prob.sum(1) - 1).abs().max().item() < 1e-4
    assert (rois[:, 1:] >= 0).all() and (rois[:, 1:] <= 511).all()
    assert (scores[:300][:-1] >= scores[:300][1:]).all() or (scores[:300] == 0).any()   # sorted kept rows (+ fillers)
    r2 = net.forward_inference(batch["data"], batch["im_info"])
    assert all(torch.equal(a, b) for a, b in zip((rois, scores, cls_prob, bbox_pred), r2))
    assert torch.equal(w0, net.P.w)
