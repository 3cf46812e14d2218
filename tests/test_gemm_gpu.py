"""tcgen05 GEMM / implicit-GEMM conv / wgrad vs a plain PyTorch fp32 reference of the same op.

Tolerance: TF32 inputs carry 10 explicit mantissa bits -> per-product relative error 2^-10; for
K-term dot products of O(1) random data |err| <~ 2^-10*sqrt(K)*|a||b|; asserted as
max|d| <= 6e-3 * sqrt(K) * rms(a) * rms(b)   (bf16: 8x that).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _tol(K, a, b, dtype):
    import torch
    base = 6e-3 * (K ** 0.5) * a.float().pow(2).mean().sqrt().item() * b.float().pow(2).mean().sqrt().item()
    return base * (8 if dtype == torch.bfloat16 else 1) + 1e-5


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 64, 64), (256, 256, 256), (300, 98, 512), (6000, 1024, 1024),
                                   (1000, 85, 1024), (20480, 256, 1024)])
def test_gemm_nt_tf32(M, N, K):
    import torch
    from sniper_b200 import ops
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda")
    b = torch.randn(N, K, device="cuda")
    c = ops.gemm_nt(a, b)
    torch.cuda.synchronize()
    ref = a.double() @ b.double().t()
    assert (c.double() - ref).abs().max().item() <= _tol(K, a, b, a.dtype)


def test_gemm_epilogue_and_bf16():
    import torch
    from sniper_b200 import ops
    torch.manual_seed(0)
    M, N, K = 512, 200, 256
    a = torch.randn(M, K, device="cuda")
    b = torch.randn(N, K, device="cuda")
    bias = torch.randn(N, device="cuda")
    scale = torch.rand(N, device="cuda") + 0.5
    res = torch.randn(M, N, device="cuda")
    c = ops.gemm_nt(a, b, scale=scale, bias=bias, residual=res, relu=True)
    ref = torch.relu((a.double() @ b.double().t()) * scale.double() + bias.double() + res.double())
    assert (c.double() - ref).abs().max().item() <= _tol(K, a, b, a.dtype) * 1.5
    c2 = ops.gemm_nt(a, b, out=c.clone(), accumulate=True)
    ref2 = ref + a.double() @ b.double().t()
    assert (c2.double() - ref2).abs().max().item() <= _tol(K, a, b, a.dtype) * 2.5
    ab, bb = a.bfloat16(), b.bfloat16()
    c3 = ops.gemm_nt(ab, bb, out_dtype=torch.float32)
    ref3 = ab.double() @ bb.double().t()
    assert (c3.double() - ref3).abs().max().item() <= 1e-2  # inputs already bf16: only fp32 accumulation error


CONVS = [
    # NB, H,  W,  Cin, Cout, k, stride, dil, pad
    (2, 32, 32, 64, 64, 1, 1, 1, 0),
    (2, 32, 32, 256, 256, 3, 1, 1, 1),
    (3, 32, 32, 512, 72, 3, 1, 2, 2),
    (2, 64, 64, 128, 128, 3, 2, 1, 1),
    (2, 64, 64, 256, 512, 1, 2, 1, 0),
    (1, 128, 128, 64, 64, 3, 1, 1, 1),
    (2, 128, 128, 128, 128, 3, 2, 1, 1),
    (1, 32, 32, 3072, 512, 3, 1, 1, 1),
]


# inference canvases: any height, widths that are multiples of 8 at the layer's stride (narrow, tall tiles; rows past Ho
# clipped by the TMA); rectangular maps; a width of 136 = 8 * 17 (tile_w 8), 88 (tile_w 8), 160 (tile_w 32), 50 rows
RECT_CONVS = [
    (2, 20, 24, 64, 64, 3, 1, 1, 1),
    (1, 50, 88, 128, 256, 1, 1, 1, 0),
    (1, 9, 136, 64, 96, 3, 1, 1, 1),
    (2, 36, 160, 64, 128, 3, 2, 1, 1),
    (1, 44, 80, 256, 512, 1, 2, 1, 0),
    (1, 22, 40, 512, 64, 3, 1, 2, 2),
]


@pytest.mark.parametrize("NB,H,W,Cin,Cout,k,stride,dil,pad", CONVS + RECT_CONVS)
def test_conv2d_nhwc_tf32(NB, H, W, Cin, Cout, k, stride, dil, pad):
    import torch
    import torch.nn.functional as F
    from sniper_b200 import ops
    torch.manual_seed(H + Cin + Cout + k)
    x = torch.randn(NB, H, W, Cin, device="cuda")
    w = torch.randn(Cout, k, k, Cin, device="cuda") * (1.0 / (k * k * Cin) ** 0.5)
    bias = torch.randn(Cout, device="cuda")
    y = ops.conv2d_nhwc(x, w.reshape(Cout, -1).contiguous(), kh=k, kw=k, stride=stride, dil=dil, pad=pad, bias=bias,
                        relu=True)
    torch.cuda.synchronize()
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.permute(0, 3, 1, 2).double(), bias.double(), stride=stride,
                   padding=pad, dilation=dil).relu().permute(0, 2, 3, 1)
    assert y.shape == ref.shape
    assert (y.double() - ref).abs().max().item() <= _tol(k * k * Cin, x, w, x.dtype)


@pytest.mark.parametrize("NB,H,W,Cin,Cout,k,stride,dil,pad", CONVS[:5])
def test_conv2d_wgrad_tf32(NB, H, W, Cin, Cout, k, stride, dil, pad):
    import torch
    import torch.nn.functional as F
    from sniper_b200 import ops
    torch.manual_seed(1 + H + Cin + Cout + k)
    x = torch.randn(NB, H, W, Cin, device="cuda")
    Ho = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
    if Cout % 32:
        Cout = 96
    dy = torch.randn(NB, Ho, Ho, Cout, device="cuda")
    dw = ops.conv2d_wgrad_nhwc(dy, x, kh=k, kw=k, stride=stride, dil=dil, pad=pad, splits=4)
    torch.cuda.synchronize()
    xd = x.permute(0, 3, 1, 2).double().requires_grad_(False)
    wd = torch.zeros(Cout, Cin, k, k, device="cuda", dtype=torch.double, requires_grad=True)
    out = F.conv2d(xd, wd, None, stride=stride, padding=pad, dilation=dil)
    out.backward(dy.permute(0, 3, 1, 2).double())
    ref = wd.grad.permute(0, 2, 3, 1).reshape(Cout, -1)
    Kred = NB * Ho * Ho
    assert (dw.double() - ref).abs().max().item() <= _tol(Kred, x, dy, x.dtype)


def _both_epilogues(fn):
    """Runs fn() with the direct (per-lane row stores) and the staged TMA-store epilogue of the tcgen05 kernel."""
    import os
    res = []
    os.environ["SNIPER_GEMM_TAIL"] = "0"      # the tail split re-associates the K sum (TMA epilogue only)
    for flag in ("0", "1"):
        os.environ["SNIPER_GEMM_TMA_STORE"] = flag
        try:
            res.append(fn())
        finally:
            os.environ.pop("SNIPER_GEMM_TMA_STORE", None)
    os.environ.pop("SNIPER_GEMM_TAIL", None)
    import torch
    torch.cuda.synchronize()
    return res


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (300, 96, 256), (6000, 1024, 512), (20480, 256, 128), (777, 512, 96)])
def test_tma_store_epilogue_matches_direct_gemm(M, N, K):
    """Same accumulators, same fp32 epilogue arithmetic -> the staged TMA-store epilogue is BIT-identical to the
    direct one: plain, scale+bias+relu, residual (separate and in place), ragged M (rows clipped by the TMA)."""
    import torch
    from sniper_b200 import ops
    torch.manual_seed(M + N)
    a = torch.randn(M, K, device="cuda")
    b = torch.randn(N, K, device="cuda")
    sc = torch.rand(N, device="cuda") + 0.5
    bi = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda")
    d, t = _both_epilogues(lambda: ops.gemm_nt(a, b))
    assert torch.equal(d, t)
    d, t = _both_epilogues(lambda: ops.gemm_nt(a, b, scale=sc, bias=bi, relu=True))
    assert torch.equal(d, t)
    d, t = _both_epilogues(lambda: ops.gemm_nt(a, b, bias=bi, residual=res))
    assert torch.equal(d, t)
    assert torch.equal(d, ops.gemm_nt(a, b, bias=bi) + res) or (d - (ops.gemm_nt(a, b, bias=bi) + res)).abs().max() < 1e-4

    def inplace():
        o = res.clone()
        ops.gemm_nt(a, b, out=o, residual=o, relu=True)
        return o
    d, t = _both_epilogues(inplace)
    assert torch.equal(d, t)
    # column-sliced output / residual views (ld > N), as the concat buffers of the network use them
    big = torch.zeros(M, N + 64, device="cuda")

    def sliced():
        o = big.clone()
        ops.gemm_nt(a, b, out=o[:, 32:32 + N], residual=res)
        return o
    d, t = _both_epilogues(sliced)
    assert torch.equal(d, t) and float(t[:, :32].abs().sum()) == 0.0 and float(t[:, 32 + N:].abs().sum()) == 0.0


@pytest.mark.parametrize("NB,H,W,Cin,Cout,k,stride,dil,pad", CONVS)
def test_tma_store_epilogue_matches_direct_conv(NB, H, W, Cin, Cout, k, stride, dil, pad):
    import torch
    from sniper_b200 import ops
    torch.manual_seed(H + Cout)
    x = torch.randn(NB, H, W, Cin, device="cuda")
    w = torch.randn(Cout, k * k * Cin, device="cuda") * 0.05
    bi = torch.randn(Cout, device="cuda")
    d, t = _both_epilogues(lambda: ops.conv2d_nhwc(x, w, kh=k, kw=k, stride=stride, dil=dil, pad=pad, bias=bi, relu=True))
    assert torch.equal(d, t)
    res = torch.randn_like(d)
    d, t = _both_epilogues(lambda: ops.conv2d_nhwc(x, w, kh=k, kw=k, stride=stride, dil=dil, pad=pad, residual=res))
    assert torch.equal(d, t)
    # accumulate=True: red.global.add vs TMA reduce-add onto the same initial values (one contribution per element)
    d, t = _both_epilogues(lambda: ops.conv2d_nhwc(x, w, kh=k, kw=k, stride=stride, dil=dil, pad=pad, out=res.clone(),
                                                   accumulate=True))
    assert torch.equal(d, t)
    # weight gradient (split-K accumulation order differs run to run -> tolerance)
    dy = torch.randn_like(d)
    if (NB, H, W, Cin, Cout, k, stride, dil, pad) in CONVS[:5] and Cout % 32 == 0:
        d, t = _both_epilogues(lambda: ops.conv2d_wgrad_nhwc(dy, x, kh=k, kw=k, stride=stride, dil=dil, pad=pad, splits=0))
        assert float((d - t).norm() / d.norm()) < 1e-5


def test_tma_store_epilogue_strided_output_map():
    """Stride-2 data gradient: the GEMM rows land on every other pixel of the (pre-zeroed / residual) output."""
    import torch
    from sniper_b200 import ops
    torch.manual_seed(2)
    NB, Ho, Wo, Cout, Cin = 2, 32, 32, 128, 256
    dy = torch.randn(NB, Ho, Wo, Cout, device="cuda")
    wt = torch.randn(Cin, Cout, device="cuda") * 0.05
    base = torch.randn(NB, 2 * Ho, 2 * Wo, Cin, device="cuda")

    def run():
        o = base.clone()
        ops.conv2d_nhwc(dy, wt, kh=1, kw=1, out=o, residual=o, out_hw=(Ho, Wo), out_map=(2 * Ho, 2 * Wo, 2, 1, 1))
        return o
    d, t = _both_epilogues(run)
    assert torch.equal(d, t)
    assert torch.equal(t[:, 0::2], base[:, 0::2]) and not torch.equal(t[:, 1::2, 1::2], base[:, 1::2, 1::2])


@pytest.mark.parametrize("tma", ["0", "1"])
def test_fused_column_statistics(tma):
    """stats= (BatchNorm statistics of the consumer fused into the epilogue): per-column sum and sum of squares of the
    STORED values, over exactly the rows that exist (ragged M, partial conv tiles), with both epilogue variants."""
    import os
    import torch
    from sniper_b200 import ops
    os.environ["SNIPER_GEMM_TMA_STORE"] = tma
    try:
        torch.manual_seed(5)
        for (M, N, K) in [(300, 96, 64), (20480, 256, 128), (130, 512, 32)]:
            a = torch.randn(M, K, device="cuda")
            b = torch.randn(N, K, device="cuda")
            bi = torch.randn(N, device="cuda")
            st = torch.zeros(2 * N, dtype=torch.float64, device="cuda")
            c = ops.gemm_nt(a, b, bias=bi, stats=st)
            ref = torch.cat([c.double().sum(0), (c.double() ** 2).sum(0)])
            assert ((st - ref).abs() / (ref.abs() + 1.0)).max().item() < 1e-5
        # conv with a partial last tile row (Ho = 20, tile = 16 x 8 pixels) and a residual
        x = torch.randn(3, 20, 16, 64, device="cuda")
        w = torch.randn(128, 9 * 64, device="cuda") * 0.05
        res = torch.randn(3, 20, 16, 128, device="cuda")
        st = torch.zeros(256, dtype=torch.float64, device="cuda")
        y = ops.conv2d_nhwc(x, w, kh=3, kw=3, pad=1, residual=res, stats=st)
        y2 = y.double().reshape(-1, 128)
        ref = torch.cat([y2.sum(0), (y2 ** 2).sum(0)])
        assert ((st - ref).abs() / (ref.abs() + 1.0)).max().item() < 1e-5
    finally:
        os.environ.pop("SNIPER_GEMM_TMA_STORE", None)


@pytest.mark.parametrize("mode", ["1", "2"])
def test_tail_split_matches_whole_tiles_and_is_deterministic(mode):
    """The last partial wave of tiles is cut into column pieces (mode 2: same K order, BIT-identical to whole tiles)
    or into K-slices whose parked accumulators are summed in slice order by the last arriving slice (mode 1: fp32
    re-association only, bit-identical run to run).  Epilogue (bias, residual, fused statistics) applied exactly once."""
    import os
    import torch
    from sniper_b200 import ops
    torch.manual_seed(8)
    cases = [(20480, 256, 2304), (6000, 128, 12544), (20480, 512, 4608), (300, 64, 8192), (20480, 1024, 256)]

    def with_mode(m, fn):
        os.environ["SNIPER_GEMM_TAIL"] = m
        try:
            return fn()
        finally:
            os.environ.pop("SNIPER_GEMM_TAIL", None)
    for (M, N, K) in cases:
        a = torch.randn(M, K, device="cuda")
        b = torch.randn(N, K, device="cuda")
        bi = torch.randn(N, device="cuda")
        res = torch.randn(M, N, device="cuda")

        def run():
            st = torch.zeros(2 * N, dtype=torch.float64, device="cuda")
            c = ops.gemm_nt(a, b, bias=bi, residual=res, stats=st)
            return c, st
        c0, s0 = with_mode("0", run)
        c1, s1 = with_mode(mode, run)
        c2, s2 = with_mode(mode, run)
        torch.cuda.synchronize()
        assert torch.equal(c1, c2)
        if mode == "2" and N % 256 == 0 and (M // 128) * (N // 256) > 148:
            assert torch.equal(c0, c1)                # column pieces
        else:
            # the K sum is re-associated (slices accumulate separately in TMEM): far below the TF32 operand rounding
            assert float((c0 - c1).abs().max()) <= 1e-4 * float(c0.abs().max())
        # statistics are those of the values actually stored, counted exactly once
        ref1 = torch.cat([c1.double().sum(0), (c1.double() ** 2).sum(0)])
        den = torch.cat([c1.double().abs().sum(0), (c1.double() ** 2).sum(0)]) + 1.0
        assert ((s1 - ref1).abs() / den).max().item() < 1e-5
    # NHWC conv (3x3, 160 tiles of 128 pixels) with the tail split on vs off
    x = torch.randn(20, 32, 32, 128, device="cuda")
    w = torch.randn(256, 9 * 128, device="cuda") * 0.05
    y0 = with_mode("0", lambda: ops.conv2d_nhwc(x, w, kh=3, kw=3, pad=1, relu=True))
    y1 = with_mode(mode, lambda: ops.conv2d_nhwc(x, w, kh=3, kw=3, pad=1, relu=True))
    if mode == "2":
        assert torch.equal(y0, y1)
    else:
        assert float((y0 - y1).abs().max()) <= 1e-4 * float(y0.abs().max())
