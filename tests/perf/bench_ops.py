"""Times the two proposal operators on the GPU (CUDA events) next to the REFERENCE's own CPU operator binaries
(oracle/_ref/libref_mpt.so / libref_mp.so, compiled from multi_proposal_target.cc / multi_proposal.cc) on the host
cores of the same box.  B = 20 chips of 512x512 (21504 anchors each), 300 rois per chip.
Usage: python tests/perf/bench_ops.py > profiles/proposal_ops_rNN.md"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402  (checker / CPU baseline only)
from sniper_b200 import ops, synth  # noqa: E402


def gpu_time(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2]


def cpu_time(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.time(); fn(); ts.append((time.time() - t0) * 1e3)
    return min(ts)


def main():
    B = 20
    cls_prob, bbox_pred, im_info, gts, vr = synth.mpt_inputs(7, B)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d = [t(x) for x in (cls_prob, bbox_pred, im_info, gts, vr)]
    rows = []
    g = gpu_time(lambda: ops.multi_proposal_target(*d))
    c = cpu_time(lambda: O.ref_multi_proposal_target(cls_prob, bbox_pred, im_info, gts, vr)) if O.ref_op_lib("libref_mpt.so") else None
    rows.append(("MultiProposalTarget fwd (training)", g, c))
    g = gpu_time(lambda: ops.multi_proposal(d[0], d[1], d[2]))
    c = cpu_time(lambda: O.ref_multi_proposal(cls_prob, bbox_pred, im_info)) if O.ref_op_lib("libref_mp.so") else None
    rows.append(("MultiProposal fwd (inference)", g, c))
    print("| operator (B = 20 chips, 21504 anchors/chip, 300 rois) | B200, this repo (ms) | reference CPU operator binary, "
          "%d host threads (ms) | ratio |" % (os.cpu_count() or 1))
    print("|---|---:|---:|---:|")
    for name, g, c in rows:
        print("| %s | %.3f | %s | %s |" % (name, g, "%.1f" % c if c else "n/a", "%.0fx" % (c / g) if c else "n/a"))
    print("\nGPU: median of 30 launches (CUDA events, inputs resident); CPU: best of 3 calls of the reference binary "
          "(its OpenMP loops are hard-wired to 8 threads).  The reference GPU build of both operators copies all scores "
          "and deltas to the host and runs this same CPU code (multi_proposal.cu:440-470).")


if __name__ == "__main__":
    main()
