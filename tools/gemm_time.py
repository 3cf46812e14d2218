"""Times single tcgen05 launches (CUDA events, rotating operands larger than L2) under different env settings.
Usage: python tools/gemm_time.py  (edit CASES / SETTINGS below)"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sniper_b200 import ops

DT = torch.bfloat16 if os.environ.get("GT_DTYPE") == "bf16" else torch.float32      # operand / output storage
CASES = [("gemm", 20480, 1024, 256), ("gemmres", 20480, 1024, 256), ("gemmr", 20480, 1024, 256), ("gemmst", 20480, 1024, 256), ("gemm", 327680, 256, 64), ("gemm", 81920, 512, 128),
         ("gemm", 81920, 128, 512), ("gemm", 20480, 256, 1024), ("conv3x3", 20, 32, 32, 256, 256),
         ("gemm", 20480, 256, 2304), ("gemm", 18944, 256, 2304), ("gemm", 37888, 256, 2304), ("conv3x3", 37, 32, 32, 256, 256)]
SETTINGS = [{"SNIPER_GEMM_2SM": "0"}, {"SNIPER_GEMM_2SM": "1"}]
if os.environ.get("GT_CASES"):          # e.g. GT_CASES="gemm:20480:1024:256;gemmr:20480:1024:256"
    CASES = [tuple([c.split(":")[0]] + [int(v) for v in c.split(":")[1:]]) for c in os.environ["GT_CASES"].split(";")]
if len(sys.argv) > 1:
    SETTINGS = [dict(kv.split("=") for kv in a.split(",") if kv) for a in sys.argv[1:]]


def timed(fn, reps=20):
    """Median-free mean over `reps` back-to-back launches replayed from a CUDA graph (no host launch floor: an eager
    ops call costs ~20 us on the host, more than the short-K kernels take)."""
    for _ in range(3):
        fn(0)
    torch.cuda.synchronize()
    if os.environ.get("GT_EAGER") != "1":
        g = torch.cuda.CUDAGraph()
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            with torch.cuda.graph(g, stream=st):
                for i in range(reps):
                    fn(i)
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return sorted(ts)[2] * 1e3 / reps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for i, (a, b) in enumerate(ev):
        a.record(); fn(i); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2] * 1e3


for case in CASES:
    if case[0] in ("gemm", "gemmres", "gemmr", "gemmst"):
        _, M, N, K = case
        nbuf = max(2, int(300e6 // (4 * (M * K + M * N))) + 1)
        A = [torch.randn(M, K, device="cuda").to(DT) for _ in range(nbuf)]
        B = torch.randn(N, K, device="cuda").to(DT)
        C = [torch.empty(M, N, device="cuda", dtype=DT) for _ in range(nbuf)]
        if case[0] != "gemm":      # gemmres: residual + BN statistics; gemmr: residual only; gemmst: statistics only
            R = [torch.randn(M, N, device="cuda").to(DT) for _ in range(nbuf)] if case[0] != "gemmst" else None
            stats_buf = torch.zeros(2 * N, dtype=torch.float64, device="cuda") if case[0] != "gemmr" else None
            fn = lambda i: ops.gemm_nt(A[i % nbuf], B, out=C[i % nbuf], residual=None if R is None else R[(i + 1) % nbuf],
                                       stats=stats_buf)
        else:
            fn = lambda i: ops.gemm_nt(A[i % nbuf], B, out=C[i % nbuf])
        flop = 2.0 * M * N * K
    else:
        _, NB, H, W, Cin, Cout = case
        nbuf = max(2, int(300e6 // (4 * NB * H * W * (Cin + Cout))) + 1)
        X = [torch.randn(NB, H, W, Cin, device="cuda").to(DT) for _ in range(nbuf)]
        Wt = (torch.randn(Cout, 9 * Cin, device="cuda") * 0.02).to(DT)
        Y = [torch.empty(NB, H, W, Cout, device="cuda", dtype=DT) for _ in range(nbuf)]
        fn = lambda i: ops.conv2d_nhwc(X[i % nbuf], Wt, kh=3, kw=3, pad=1, out=Y[i % nbuf])
        flop = 2.0 * NB * H * W * Cout * 9 * Cin
    out = []
    for st in SETTINGS:
        for k in ("SNIPER_GEMM_STG", "SNIPER_GEMM_TAIL", "SNIPER_GEMM_TAIL_MAXS", "SNIPER_GEMM_TAIL_MAXP", "SNIPER_GEMM_BN", "SNIPER_GEMM_2SM", "SNIPER_GEMM_TMA_STORE", "SNIPER_GEMM_SPEC"):
            os.environ.pop(k, None)
        os.environ.update(st)
        us = timed(fn)
        out.append("%s: %.1f us %.0f TF/s" % (",".join("%s=%s" % kv for kv in st.items()).replace("SNIPER_GEMM_", ""), us, flop / us / 1e6))
    print(case, " | ".join(out), flush=True)
    del fn
