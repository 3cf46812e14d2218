"""One tcgen05 shape, a few launches: the command ncu wraps for a --set full capture of the dominant kernel.
usage: python tools/gemm_one.py M N K [bf16] [res] [stats]"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sniper_b200 import ops

M, N, K = map(int, sys.argv[1:4])
flags = sys.argv[4:]
dt = torch.bfloat16 if "bf16" in flags else torch.float32
a = torch.randn(M, K, device="cuda").to(dt)
b = torch.randn(N, K, device="cuda").to(dt)
c = torch.empty(M, N, device="cuda", dtype=dt)
r = torch.randn(M, N, device="cuda").to(dt) if "res" in flags else None
st = torch.zeros(2 * N, dtype=torch.float64, device="cuda") if "stats" in flags else None
for _ in range(3):
    ops.gemm_nt(a, b, out=c, residual=r, stats=st)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(2):
    ops.gemm_nt(a, b, out=c, residual=r, stats=st)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
