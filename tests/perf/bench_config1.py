"""BASELINE.json configs[0] (CPU plumbing, no GPU): lib/chips chip generation over the three SNIPER scales of one
synthetic 1333x800 image (20 GT boxes + 2000 proposal boxes), then cpu_nms (thresh 0.7) and cpu_soft_nms (sigma 0.55) on
6000 detections.  Times this repo's host entry points (libsniper_b200.so, host_ops.cpp) next to the reference's own code
built into oracle/_ref (cchips.cpp compiled as it lies; cpu_nms.pyx cythonized) and checks that the outputs are identical.
Usage: python tests/perf/bench_config1.py > profiles/config1_host_rNN.md"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
import oracle_lib as O  # noqa: E402  (reference binaries live behind it)
from sniper_b200 import host  # noqa: E402


def boxes_like(seed, n, W=1333, H=800):
    rng = np.random.RandomState(seed)
    s = np.exp(rng.uniform(np.log(8), np.log(400), n))
    ar = np.exp(rng.uniform(np.log(0.5), np.log(2.0), n))
    w, h = s * np.sqrt(ar), s / np.sqrt(ar)
    cx, cy = rng.uniform(0, W, n), rng.uniform(0, H, n)
    return np.stack([np.clip(cx - w / 2, 0, W - 1), np.clip(cy - h / 2, 0, H - 1), np.clip(cx + w / 2, 0, W - 1),
                     np.clip(cy + h / 2, 0, H - 1)], 1).astype(np.float32)


def best(fn, reps=7):
    ts = []
    out = None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), out


def main():
    rows = []
    boxes0 = np.concatenate([boxes_like(0, 20), boxes_like(1, 2000)])
    for scale in (3.0, 1.667, 0.384):
        W, H = int(1333 * scale), int(800 * scale)
        b = boxes0 * np.float32(scale)
        b[:, [0, 2]] = np.clip(b[:, [0, 2]], 0, W - 1)
        b[:, [1, 3]] = np.clip(b[:, [1, 3]], 0, H - 1)

        def ours():
            host.srand(1)
            return host.chips_generate(b, W, H, 512, 58)
        t_o, c_o = best(ours)
        if O.ref_chips() is not None:
            t_r, c_r = best(lambda: O.ref_chips_generate(b, W, H, 512, 58, seed=1))
            same = np.array_equal(c_o, c_r)
        else:
            t_r, same = float("nan"), None
        rows.append(("chips generate, scale %.3f (%dx%d, %d boxes -> %d chips)" % (scale, W, H, len(b), len(c_o)), t_o, t_r, same))
    rng = np.random.RandomState(2)
    d = boxes_like(3, 6000)
    dets = np.concatenate([d, rng.permutation(6000).reshape(-1, 1) / 6000.0], 1).astype(np.float32)
    try:
        import ref_cpu_nms
    except ImportError:
        ref_cpu_nms = None
    t_o, k_o = best(lambda: host.cpu_nms(dets, 0.7), 3)
    if ref_cpu_nms:
        t_r, k_r = best(lambda: [int(i) for i in ref_cpu_nms.cpu_nms(dets.copy(), 0.7)], 3)
        rows.append(("cpu_nms, 6000 dets, thresh 0.7 (%d kept)" % len(k_o), t_o, t_r, k_o == k_r))
    t_o, s_o = best(lambda: host.cpu_soft_nms(dets.copy(), sigma=0.55, Nt=0.3, threshold=0.001, method=2), 3)
    if ref_cpu_nms:
        t_r, s_r = best(lambda: np.array(ref_cpu_nms.cpu_soft_nms(dets.copy(), sigma=0.55, Nt=0.3, threshold=0.001, method=2)), 3)
        same = s_o.shape == s_r.shape and np.array_equal(s_o[:, :4], s_r[:, :4]) and np.abs(s_o[:, 4] - s_r[:, 4]).max() <= 2e-7
        rows.append(("cpu_soft_nms, 6000 dets, sigma 0.55 (%d rows survive)" % len(s_o), t_o, t_r, bool(same)))
    cpu = "unknown"
    for line in open("/proc/cpuinfo"):
        if line.startswith("model name"):
            cpu = line.split(":", 1)[1].strip()
            break
    print("# BASELINE config 1 (host plumbing), %s, single thread\n" % cpu)
    print("| step | this repo, host_ops.cpp (ms) | reference's own code in oracle/_ref (ms) | identical output |")
    print("|---|---:|---:|---|")
    for name, a, b_, same in rows:
        print("| %s | %.3f | %.3f | %s |" % (name, a, b_, same))
    print("\nBest of 3-7 calls; `srand(1)` before every chip generation on both sides (same libstdc++ `rand()` stream).")


if __name__ == "__main__":
    main()
