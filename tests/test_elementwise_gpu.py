"""HBM-bound layers and losses vs plain PyTorch fp32/fp64 references of the same op."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_bn_train_fwd_bwd_matches_torch():
    import torch
    from sniper_b200 import ops
    torch.manual_seed(0)
    M, C = 4096 + 40, 256
    x = (torch.randn(M, C, device="cuda") * 2 + 0.5)
    gamma = torch.rand(C, device="cuda") + 0.5
    beta = torch.randn(C, device="cuda")
    bn = ops.BNState(C, "cuda", gamma.clone(), beta.clone(), torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda"))
    ops.bn_stats(x, bn, eps=2e-5, momentum=0.9)
    y = ops.affine_act(x, bn.scale, bn.shift, relu=True)
    dy = torch.randn(M, C, device="cuda")
    add = torch.randn(M, C, device="cuda")
    dx = ops.bn_relu_bwd(x, dy, bn, add=add)
    xr = x.double().requires_grad_(True)
    g, b = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rm, rv = torch.zeros(C, device="cuda", dtype=torch.double), torch.ones(C, device="cuda", dtype=torch.double)
    yr = torch.relu(torch.nn.functional.batch_norm(xr, rm, rv, g, b, True, 0.1, 2e-5))
    yr.backward(dy.double())
    assert (y.double() - yr).abs().max().item() < 1e-4
    assert (dx.double() - (xr.grad + add.double())).abs().max().item() < 1e-4
    assert (bn.dgamma.double() - g.grad).abs().max().item() < 2e-2 * g.grad.abs().max().item() * 0.05 + 1e-2
    assert (bn.dbeta.double() - b.grad).abs().max().item() < 1e-2
    assert (bn.moving_mean.double() - rm).abs().max().item() < 1e-5
    assert (bn.moving_var.double() - rv).abs().max().item() < 1e-4
    assert float(bn.sums.abs().sum()) == 0.0  # scratch left zeroed


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_bn_apply_train_equals_finalize_then_apply(dtype):
    """sniper_bn_apply_train (finalisation folded into the apply pass) == sniper_bn_finalize + sniper_affine_act, bit for
    bit, including the published mean / invstd / scale / shift and the moving statistics; the forward accumulator is
    cleared by the batched parameter-gradient launch."""
    import torch
    from sniper_b200 import ops
    torch.manual_seed(3)
    dt = getattr(torch, dtype)
    M, C = 20480 + 24, 256
    x = (torch.randn(M, C, device="cuda") * 1.5 + 0.3).to(dt)
    a, b = ops.BNState(C, "cuda"), ops.BNState(C, "cuda")
    for st in (a, b):
        st.gamma.copy_(torch.linspace(0.5, 1.5, C)); st.beta.copy_(torch.linspace(-1, 1, C))
        st.moving_mean.fill_(0.25); st.moving_var.fill_(2.0)
    xs = x.double()
    sums = torch.cat([xs.sum(0), (xs * xs).sum(0)])
    a.sums.copy_(sums); b.sums_f.copy_(sums)
    ops.bn_finalize(a, M, eps=2e-5, momentum=0.9)
    ya = ops.affine_act(x, a.scale, a.shift, relu=True)
    yb = ops.bn_apply_train(x, b, eps=2e-5, momentum=0.9, relu=True)
    assert torch.equal(ya, yb)
    for f in ("mean", "invstd", "scale", "shift", "moving_mean", "moving_var"):
        assert torch.equal(getattr(a, f), getattr(b, f)), f
    assert torch.equal(b.sums_f, sums)                      # untouched by the apply pass
    b.dgamma = torch.zeros(C, device="cuda"); b.dbeta = torch.zeros(C, device="cuda")
    ops.bn_param_grad_batched(ops.bn_param_grad_jobs([b], "cuda"))
    assert float(b.sums_f.abs().sum()) == 0.0


def test_strided_rows_and_frozen_bn():
    import torch
    from sniper_b200 import ops
    torch.manual_seed(1)
    buf = torch.randn(2, 8, 8, 96, device="cuda")
    x = buf[..., 32:96]          # channel slice: ld = 96
    bn = ops.BNState(64, "cuda")
    bn.moving_mean.normal_(); bn.moving_var.uniform_(0.5, 2); bn.gamma.uniform_(0.5, 1.5); bn.beta.normal_()
    ops.bn_frozen(bn, eps=2e-5)
    y = ops.affine_act(x, bn.scale, bn.shift, relu=True)
    ref = torch.relu((x - bn.moving_mean) / torch.sqrt(bn.moving_var + 2e-5) * bn.gamma + bn.beta)
    assert (y - ref).abs().max().item() < 1e-5


def test_maxpool_stem_transpose_sgd():
    import torch
    import torch.nn.functional as F
    from sniper_b200 import ops
    torch.manual_seed(2)
    x = torch.randn(2, 18, 22, 64, device="cuda")
    y = ops.maxpool3x3s2(x)
    ref = F.max_pool2d(x.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(y, ref)
    img = torch.randn(2, 3, 64, 96, device="cuda") * 50
    w = torch.randn(64, 7, 7, 3, device="cuda") * 0.05
    isc, ish = torch.rand(3, device="cuda") + 0.5, torch.randn(3, device="cuda")
    osc, osh = torch.rand(64, device="cuda") + 0.5, torch.randn(64, device="cuda")
    o = ops.stem_conv(img, w, isc, ish, osc, osh)
    xin = img * isc.view(1, 3, 1, 1) + ish.view(1, 3, 1, 1)
    r = F.conv2d(xin.double(), w.permute(0, 3, 1, 2).double(), None, stride=2, padding=3)
    r = torch.relu(r * osc.double().view(1, 64, 1, 1) + osh.double().view(1, 64, 1, 1)).permute(0, 2, 3, 1)
    assert o.shape == r.shape and (o.double() - r).abs().max().item() < 1e-3
    wt = torch.randn(96, 9, 64, device="cuda")
    sel = torch.tensor([8, 7, 6, 5, 4, 3, 2, 1, 0], dtype=torch.int32, device="cuda")
    t = ops.weight_transpose(wt, 96, 9, 64, sel).view(64, 9, 96)
    assert torch.equal(t, wt.flip(1).permute(2, 1, 0).contiguous())
    n = 1001
    wv, mom, g = torch.randn(n + 3, device="cuda")[:n], torch.randn(n + 3, device="cuda")[:n], torch.randn(n + 3, device="cuda")[:n]
    w0, m0 = wv.clone(), mom.clone()
    ops.sgd_mom(wv, mom, g, lr=0.01, wd=1e-4, momentum=0.9)
    m1 = 0.9 * m0 - 0.01 * 1e-4 * w0 - 0.01 * g
    assert (mom - m1).abs().max().item() < 1e-6 and (wv - (w0 + m1)).abs().max().item() < 1e-6


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_stem_on_tensor_cores(dtype):
    """im2col of bn_data(x) (bit-exact against its definition, odd sizes and a ragged last block of columns) + tcgen05
    GEMM with bn0 + ReLU: equals the float64 convolution of the operands the MMA really reads (fp32 words truncated to
    TF32, or the stored bf16 values) to accumulation rounding, and the FP32-FMA stem kernel to TF32 / bf16 precision."""
    import torch
    import torch.nn.functional as F
    from sniper_b200 import ops
    torch.manual_seed(4)
    dt = getattr(torch, dtype)
    NB, H, W = 2, 70, 200                      # Ho 35, Wo 100: one full block of 64 columns + a ragged one of 36
    img = torch.randn(NB, 3, H, W, device="cuda") * 50
    w = torch.randn(64, 7, 7, 3, device="cuda") * 0.05
    isc, ish = torch.rand(3, device="cuda") / 50 + 0.01, torch.randn(3, device="cuda") * 0.1
    osc, osh = torch.rand(64, device="cuda") + 0.5, torch.randn(64, device="cuda")
    rows = ops.stem_rows(w, dt)
    Kp = rows.shape[1]
    assert Kp == (160 if dt == torch.float32 else 192) and float(rows[:, 147:].abs().max()) == 0.0
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    col = torch.empty(NB * Ho * Wo, Kp, device="cuda", dtype=dt)
    ops.check(ops.lib().sniper_stem_im2col(img.data_ptr(), isc.data_ptr(), ish.data_ptr(), col.data_ptr(), NB, H, W, Kp,
                                           0 if dt == torch.float32 else 1, torch.cuda.current_stream().cuda_stream))
    xin = torch.addcmul(ish.view(1, 3, 1, 1), img, isc.view(1, 3, 1, 1))          # one fma, like the kernel
    ref = F.unfold(xin, 7, padding=3, stride=2).view(NB, 3, 49, Ho * Wo).permute(0, 3, 2, 1).reshape(NB * Ho * Wo, 147)
    assert torch.equal(col[:, :147], ref.to(dt)) and float(col[:, 147:].abs().max()) == 0.0
    y = ops.stem_conv_tc(img, rows, isc, ish, osc, osh, out_dtype=dt)
    if dt == torch.float32:
        trunc = lambda t: (t.contiguous().view(torch.int32) & -8192).view(torch.float32)
        a, b = trunc(col), trunc(rows)
    else:
        a, b = col, rows
    r = torch.relu((a.double() @ b.double().t()) * osc.double() + osh.double()).view(NB, Ho, Wo, 64)
    tol = 2e-5 if dt == torch.float32 else 1e-2            # accumulation order / the bf16 rounding of the stored output
    assert y.shape == r.shape and float((y.double() - r).abs().max()) <= tol * float(r.abs().max())
    fma = ops.stem_conv(img, w, isc, ish, osc, osh)
    assert float((y.float() - fma).abs().max()) <= (2e-3 if dt == torch.float32 else 2e-2) * float(fma.abs().max())


def test_losses_match_torch():
    import torch
    import torch.nn.functional as F
    from sniper_b200 import ops
    torch.manual_seed(3)
    B, H, W, A = 3, 8, 8, 21
    score = torch.randn(B, H, W, 128, device="cuda")
    label = torch.randint(-1, 2, (B, A * H * W), device="cuda").float()
    cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.count_valid(label, cnt)
    assert int(cnt) == int((label != -1).sum())
    prob = torch.zeros(B, H, W, 42, device="cuda")
    ds = torch.zeros(B, H, W, 128, device="cuda")
    loss = torch.zeros(1, device="cuda")
    ops.rpn_softmax_loss(score, label, A, 1.0, cnt, prob, ds, loss)
    s = score[..., :42].double().requires_grad_(True)
    lg = torch.stack([s[..., :A], s[..., A:]], -1)                       # [B,H,W,A,2]
    lab = label.view(B, A, H, W).permute(0, 2, 3, 1).long()              # (a,h,w) -> [B,H,W,A]
    ce = F.cross_entropy(lg.reshape(-1, 2), lab.reshape(-1), ignore_index=-1, reduction="sum")
    (ce / max(1, int(cnt))).backward()
    assert (ds[..., :42].double() - s.grad).abs().max().item() < 1e-6
    assert abs(float(loss) - float(ce)) < 1e-2
    p = torch.softmax(lg, -1)
    assert (prob[..., :A].double() - p[..., 0]).abs().max().item() < 1e-6
    # rpn smooth l1
    pred = torch.randn(B, H, W, 128, device="cuda") * 2
    tgt = torch.randn(B, 4 * A, H, W, device="cuda")
    wgt = (torch.rand(B, 4 * A, H, W, device="cuda") > 0.7).float()
    dp = torch.zeros(B, H, W, 128, device="cuda")
    l2 = torch.zeros(1, device="cuda")
    ops.rpn_smooth_l1_loss(pred, tgt, wgt, 4 * A, 0.25, dp, l2)
    pr = pred[..., :84].double().requires_grad_(True)
    d = pr - tgt.permute(0, 2, 3, 1).double()
    sl = (wgt.permute(0, 2, 3, 1).double() * F.smooth_l1_loss(d, torch.zeros_like(d), reduction="none", beta=1.0)).sum()
    (sl * 0.25).backward()
    assert (dp[..., :84].double() - pr.grad).abs().max().item() < 1e-6
    assert abs(float(l2) - float(sl)) < 1e-2
    # rcnn softmax / smooth l1
    N, K = 700, 81
    lg2 = torch.randn(N, 88, device="cuda")
    lb = torch.randint(0, K, (N,), device="cuda").float()
    lb[::17] = -1
    c2 = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.count_valid(lb, c2)
    pb, gr, l3 = torch.zeros(N, K, device="cuda"), torch.zeros(N, 88, device="cuda"), torch.zeros(1, device="cuda")
    ops.softmax_ce(lg2, lb, K, 1.0, c2, pb, gr, l3)
    z = lg2[:, :K].double().requires_grad_(True)
    ce2 = F.cross_entropy(z, lb.long(), ignore_index=-1, reduction="sum")
    (ce2 / int(c2)).backward()
    assert (gr[:, :K].double() - z.grad).abs().max().item() < 1e-6
    assert abs(float(l3) - float(ce2)) < 5e-2
    bp = torch.randn(N, 88, device="cuda")
    bt, bw = torch.randn(N, 4, device="cuda"), (torch.rand(N, 4, device="cuda") > 0.5).float()
    g4, l4 = torch.zeros(N, 88, device="cuda"), torch.zeros(1, device="cuda")
    ops.smooth_l1_loss(bp[:, 81:85], bt, bw, 4, 1.0 / 3008, g4[:, 81:85], l4)
    q = bp[:, 81:85].double().requires_grad_(True)
    sl2 = (bw.double() * F.smooth_l1_loss(q - bt.double(), torch.zeros(N, 4, device="cuda", dtype=torch.double), reduction="none")).sum()
    (sl2 / 3008).backward()
    assert (g4[:, 81:85].double() - q.grad).abs().max().item() < 1e-7


@pytest.mark.gpu
def test_batched_weight_transpose_and_bn_param_grad():
    """One-launch-per-step variants == the per-layer entry points."""
    import torch
    from sniper_b200 import ops
    torch.manual_seed(4)
    jobs, refs = [], []
    for (Cout, T, Cin, sel) in [(96, 9, 64, [8, 7, 6, 5, 4, 3, 2, 1, 0]), (256, 1, 1024, [0]), (512, 9, 512, [0, 2, 6, 8]),
                                (72, 9, 40, [4])]:
        w = torch.randn(Cout, T * Cin, device="cuda")
        s = torch.tensor(sel, dtype=torch.int32, device="cuda")
        out = torch.zeros(Cin, len(sel) * Cout, device="cuda")
        jobs.append((w, out, s, Cout, T, Cin))
        refs.append(ops.weight_transpose(w, Cout, T, Cin, s))
    table = ops.weight_transpose_jobs(jobs, "cuda")
    ops.weight_transpose_batched(table)
    for (_, out, *_), ref in zip(jobs, refs):
        assert torch.equal(out, ref)
    states = []
    for C in (64, 256, 1000):
        b = ops.BNState(C, "cuda", dgamma=torch.randn(C, device="cuda"), dbeta=torch.randn(C, device="cuda"))
        b.sums.copy_(torch.randn(2 * C, dtype=torch.float64, device="cuda"))
        states.append((b, b.dgamma.clone(), b.dbeta.clone(), b.sums.clone()))
    ops.bn_param_grad_batched(ops.bn_param_grad_jobs([t[0] for t in states], "cuda"))
    for b, g0, b0, s0 in states:
        assert torch.equal(b.dbeta, b0 + s0[:b.C].float()) and torch.equal(b.dgamma, g0 + s0[b.C:].float())
        assert float(b.sums.abs().sum()) == 0.0

