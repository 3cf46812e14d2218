"""Runs each tcgen05 contraction case in its own subprocess with a timeout (a wrong mbarrier/descriptor
can hang the kernel) and prints max-abs error vs fp64.  Usage: python tools/gemm_probe.py [case ...]"""
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {
    "g_128_128_32": "gemm 128 128 32",
    "g_128_64_64": "gemm 128 64 64",
    "g_128_128_128": "gemm 128 128 128",
    "g_256_256_256": "gemm 256 256 256",
    "g_300_98_512": "gemm 300 98 512",
    "g_6000_1024_1024": "gemm 6000 1024 1024",
    "g_bf16_256_128_128": "gemmbf 256 128 128",
    "c_1x1": "conv 2 32 32 64 64 1 1 1 0",
    "c_3x3": "conv 2 32 32 256 256 3 1 1 1",
    "c_3x3_s2": "conv 2 64 64 128 128 3 2 1 1",
    "c_3x3_d2": "conv 2 32 32 512 96 3 1 2 2",
    "c_3x3_128": "conv 1 128 128 64 64 3 1 1 1",
    "w_1x1": "wgrad 2 32 32 64 64 1 1 1 0",
    "w_3x3": "wgrad 2 32 32 128 128 3 1 1 1",
    "w_3x3_s2": "wgrad 2 64 64 128 128 3 2 1 1",
}

CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from sniper_b200 import ops
import torch.nn.functional as F
a = sys.argv[1:]
kind = a[0]
torch.manual_seed(0)
if kind in ("gemm", "gemmbf"):
    M, N, K = map(int, a[1:4])
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda")
    if kind == "gemmbf": A, B = A.bfloat16(), B.bfloat16()
    C = ops.gemm_nt(A, B); torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    err = (C.double() - ref).abs()
    print("maxerr %%.3e  ref_rms %%.3e  bad_frac %%.4f" %% (err.max().item(), ref.pow(2).mean().sqrt().item(), (err > 0.05 * K ** 0.5).double().mean().item()))
else:
    NB, H, W, Cin, Cout, k, s, d, p = map(int, a[1:10])
    x = torch.randn(NB, H, W, Cin, device="cuda")
    Ho = (H + 2 * p - d * (k - 1) - 1) // s + 1
    if kind == "conv":
        w = torch.randn(Cout, k, k, Cin, device="cuda") / (k * k * Cin) ** 0.5
        y = ops.conv2d_nhwc(x, w.reshape(Cout, -1).contiguous(), kh=k, kw=k, stride=s, dil=d, pad=p); torch.cuda.synchronize()
        ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.permute(0, 3, 1, 2).double(), None, stride=s, padding=p, dilation=d).permute(0, 2, 3, 1)
        err = (y.double() - ref).abs()
        print("maxerr %%.3e  ref_rms %%.3e  bad_frac %%.4f" %% (err.max().item(), ref.pow(2).mean().sqrt().item(), (err > 0.05).double().mean().item()))
    else:
        dy = torch.randn(NB, Ho, Ho, Cout, device="cuda")
        dw = ops.conv2d_wgrad_nhwc(dy, x, kh=k, kw=k, stride=s, dil=d, pad=p, splits=4); torch.cuda.synchronize()
        wd = torch.zeros(Cout, Cin, k, k, device="cuda", dtype=torch.double, requires_grad=True)
        F.conv2d(x.permute(0, 3, 1, 2).double(), wd, None, stride=s, padding=p, dilation=d).backward(dy.permute(0, 3, 1, 2).double())
        ref = wd.grad.permute(0, 2, 3, 1).reshape(Cout, -1)
        err = (dw.double() - ref).abs()
        print("maxerr %%.3e  ref_rms %%.3e  bad_frac %%.4f" %% (err.max().item(), ref.pow(2).mean().sqrt().item(), (err > 0.02 * ref.pow(2).mean().sqrt().item() + 0.05).double().mean().item()))
''' % ROOT

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for n in names:
        try:
            r = subprocess.run([sys.executable, "-c", CHILD] + CASES[n].split(), capture_output=True, text=True, timeout=90)
            out = (r.stdout.strip() or r.stderr.strip()[-600:])
            print("%-22s rc=%d %s" % (n, r.returncode, out), flush=True)
        except subprocess.TimeoutExpired:
            print("%-22s TIMEOUT (hang)" % n, flush=True)
