"""Mixed-precision (bf16 storage, fp32 accumulation / arithmetic) variants of the hot-path kernels against PyTorch
float64 on the SAME bf16-rounded inputs.  Tolerances: the result is rounded once to bf16 on store (relative 2^-9 =
1.95e-3 of its magnitude) + fp32 accumulation noise; statistics / fp32 outputs only carry accumulation noise."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
BF = 2.0 ** -8          # bf16 keeps 8 significant bits: half an ulp of x is 2^(floor(log2|x|) - 8) <= |x| * 2^-8


def _close_bf16(got, ref, extra=0.0):
    """|got - ref| <= half an ulp of bf16 at |ref| (+ slack for double rounding) + extra"""
    import torch
    err = (got.double() - ref).abs()
    bound = ref.abs() * BF * 1.01 + 1e-6 + extra
    assert (err <= bound).all(), (err - bound).max().item()


@pytest.mark.parametrize("M,N,K,res,relu,stats", [(512, 256, 256, True, True, True), (300, 128, 1024, False, False, True),
                                                 (1000, 1024, 64, True, False, False), (128, 64, 128, True, True, True)])
def test_gemm_bf16_tma_epilogue(M, N, K, res, relu, stats):
    import torch
    from sniper_b200 import ops
    torch.manual_seed(M + N)
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    scale = torch.rand(N, device="cuda") + 0.5
    r = torch.randn(M, N, device="cuda").bfloat16() if res else None
    st = torch.zeros(2 * N, dtype=torch.float64, device="cuda") if stats else None
    c = ops.gemm_nt(a, b, scale=scale, bias=bias, residual=r, relu=relu, stats=st)
    assert c.dtype == torch.bfloat16
    ref = (a.double() @ b.double().t()) * scale.double() + bias.double()
    if res:
        ref = ref + r.double()
    if relu:
        ref = torch.relu(ref)
    _close_bf16(c, ref, extra=2e-5 * K ** 0.5)
    if stats:      # statistics of the ROUNDED values (what the consumer reads)
        cd = c.double()
        assert (st[:N] - cd.sum(0)).abs().max().item() <= 1e-4 * M ** 0.5 + 1e-5 * cd.abs().sum(0).max().item()
        assert ((st[N:] - (cd * cd).sum(0)).abs() / ((cd * cd).sum(0) + 1)).max().item() <= 1e-5
    c32 = ops.gemm_nt(a, b, out_dtype=torch.float32)
    assert (c32.double() - a.double() @ b.double().t()).abs().max().item() <= 2e-5 * K ** 0.5 + 1e-5


@pytest.mark.parametrize("NB,H,Cin,Cout,k,stride,dil,pad", [(2, 32, 256, 1024, 1, 1, 1, 0), (2, 32, 256, 256, 3, 1, 1, 1),
                                                          (2, 64, 128, 128, 3, 2, 1, 1), (1, 32, 512, 128, 3, 1, 2, 2),
                                                          (2, 128, 64, 64, 1, 1, 1, 0), (2, 64, 512, 1024, 1, 2, 1, 0)])
def test_conv_bf16_forward_dgrad_wgrad(NB, H, Cin, Cout, k, stride, dil, pad):
    import torch
    import torch.nn.functional as F
    from sniper_b200 import model
    P = model.ParamStore()
    c = model.Conv(P, "c", Cin, Cout, k, stride, dil, pad, lowp=True)
    P.finalize("cuda", lowp=True)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    c.init(device="cuda", gen=g)
    P.sync_lowp()
    torch.manual_seed(0)
    x = torch.randn(NB, H, H, Cin, device="cuda").bfloat16()
    st = torch.zeros(2 * Cout, dtype=torch.float64, device="cuda")
    y = c.fwd(x, stats=st if Cout % 32 == 0 else None)
    assert y.dtype == torch.bfloat16 and c.w.dtype == torch.bfloat16
    wd = c.w.double().view(Cout, k, k, Cin).permute(0, 3, 1, 2).requires_grad_(True)
    xd = x.double().permute(0, 3, 1, 2).requires_grad_(True)
    yr = F.conv2d(xd, wd, None, stride=stride, padding=pad, dilation=dil)
    _close_bf16(y, yr.permute(0, 2, 3, 1).detach(), extra=2e-5 * (k * k * Cin) ** 0.5)
    if Cout % 32 == 0:
        yd = y.double().reshape(-1, Cout)
        assert ((st[:Cout] - yd.sum(0)).abs() / (yd.abs().sum(0) + 1)).max().item() <= 1e-5
    dy = torch.randn_like(y)
    yr.backward(dy.double().permute(0, 3, 1, 2))
    c.prepare_bwd()
    dx = c.bwd_data(dy, (H, H))
    assert dx.dtype == torch.bfloat16
    _close_bf16(dx, xd.grad.permute(0, 2, 3, 1), extra=3e-5 * (k * k * Cout) ** 0.5 * dy.double().abs().max().item())
    c.bwd_weight(dy, x, 4)
    gw = P.grad("c_weight").view(Cout, k, k, Cin).double()
    ref = wd.grad.permute(0, 2, 3, 1)
    assert ((gw - ref).norm() / ref.norm()).item() < 1e-4          # bf16 operands are exact inputs here; fp32 accumulation
    # residual path
    r = torch.randn(NB, H, H, Cin, device="cuda").bfloat16()
    dx2 = c.bwd_data(dy, (H, H), out=r.clone(), residual=r)
    _close_bf16(dx2, xd.grad.permute(0, 2, 3, 1) + r.double(), extra=3e-5 * (k * k * Cout) ** 0.5 * dy.double().abs().max().item())


def test_elementwise_bf16():
    import torch
    from sniper_b200 import ops
    torch.manual_seed(2)
    M, C = 4096, 256
    x = (torch.randn(M, C, device="cuda") * 2 + 0.3).bfloat16()
    bn = ops.BNState(C, "cuda", dgamma=torch.zeros(C, device="cuda"), dbeta=torch.zeros(C, device="cuda"))
    bn.gamma.uniform_(0.5, 1.5); bn.beta.normal_(0, 0.2)
    ops.bn_stats(x, bn, eps=2e-5, momentum=0.9)
    xd = x.double()
    mean, var = xd.mean(0), xd.var(0, unbiased=False)
    assert (bn.mean.double() - mean).abs().max().item() < 1e-5
    assert ((bn.invstd.double() - 1 / (var + 2e-5).sqrt()).abs() * (var + 2e-5).sqrt()).max().item() < 1e-5
    y = ops.affine_act(x, bn.scale, bn.shift, relu=True)
    assert y.dtype == torch.bfloat16
    yr = torch.relu((xd - mean) / (var + 2e-5).sqrt() * bn.gamma.double() + bn.beta.double())
    _close_bf16(y, yr, extra=1e-5)
    dy = torch.randn(M, C, device="cuda").bfloat16()
    add = torch.randn(M, C, device="cuda").bfloat16()
    dx = ops.bn_relu_bwd(x, dy, bn, add=add)
    xr = xd.clone().requires_grad_(True)
    g = bn.gamma.double().clone().requires_grad_(True)
    b = bn.beta.double().clone().requires_grad_(True)
    yy = torch.relu(torch.nn.functional.batch_norm(xr, None, None, g, b, True, 0.0, 2e-5))
    yy.backward(dy.double())
    _close_bf16(dx, xr.grad + add.double(), extra=2e-5)
    assert ((bn.dgamma.double() - g.grad).abs().max() / g.grad.abs().max()).item() < 1e-5
    assert ((bn.dbeta.double() - b.grad).abs().max() / b.grad.abs().max()).item() < 1e-5
    # relu_bwd, maxpool, casts
    act = torch.relu(torch.randn(M, C, device="cuda")).bfloat16()
    rb = ops.relu_bwd(act, dy)
    assert torch.equal(rb, torch.where(act > 0, dy, torch.zeros_like(dy)))
    img = torch.randn(2, 16, 16, 64, device="cuda").bfloat16()
    mp = ops.maxpool3x3s2(img)
    mr = torch.nn.functional.max_pool2d(img.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(mp.float(), mr)
    cat = torch.zeros(M, 3 * C, device="cuda")
    ops.cast_rows(x, out=cat[:, C:2 * C])
    assert torch.equal(cat[:, C:2 * C], x.float()) and float(cat[:, :C].abs().sum()) == 0
    back = ops.cast_rows(cat[:, C:2 * C], dtype=torch.bfloat16)
    assert torch.equal(back, x)


def test_stem_and_deform_im2col_bf16():
    import torch
    from sniper_b200 import ops
    torch.manual_seed(3)
    data = torch.randn(2, 3, 64, 64, device="cuda") * 60
    w = torch.randn(64, 7, 7, 3, device="cuda") * 0.1
    one, zero = torch.ones(3, device="cuda") / 60, torch.zeros(3, device="cuda")
    os_, ot = torch.rand(64, device="cuda") + 0.5, torch.randn(64, device="cuda") * 0.1
    y32 = ops.stem_conv(data, w, one, zero, os_, ot)
    y16 = ops.stem_conv(data, w, one, zero, os_, ot, out_dtype=torch.bfloat16)
    assert torch.equal(y16, y32.bfloat16())
    x = torch.randn(1, 10, 10, 512, device="cuda").bfloat16()
    off = torch.zeros(1, 10, 10, 96, device="cuda")
    off[..., :72] = torch.randn(1, 10, 10, 72, device="cuda") * 1.5
    col16 = ops.deform_im2col(x, off, kh=3, kw=3, stride=1, dil=2, pad=2, dgroups=4)
    col32 = ops.deform_im2col(x.float(), off, kh=3, kw=3, stride=1, dil=2, pad=2, dgroups=4)
    assert col16.dtype == torch.bfloat16 and torch.equal(col16, col32.bfloat16())
    dcol = torch.randn_like(col32).bfloat16()
    dx16, do16 = ops.deform_col2im(dcol, x, off, kh=3, kw=3, stride=1, dil=2, pad=2, dgroups=4)
    dx32, do32 = ops.deform_col2im(dcol.float(), x.float(), off, kh=3, kw=3, stride=1, dil=2, pad=2, dgroups=4)
    assert (dx16 - dx32).abs().max().item() <= 1e-4 * dx32.abs().max().item()      # same fp32 math, atomics order only
    assert (do16 - do32).abs().max().item() <= 1e-4 * do32.abs().max().item()
