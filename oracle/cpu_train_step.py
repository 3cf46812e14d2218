"""ORACLE / CPU baseline -- test & measurement infrastructure only (never on the product path).

Times SNIPER training steps (ResNet-101 Faster-R-CNN/R-FCN, 512x512 chips, fp32) on the host cores: the reference's
MXNet CPU stack cannot be built offline, so the graph of oracle/torch_graph.py (the same float64
restatement the parity test uses, here in float32) runs on PyTorch-CPU: dense layers through oneDNN (stand-in for
MXNet's im2col+OpenBLAS), the deformable convolution as a real bilinear gather + GEMM, MultiProposalTarget through the
reference's OWN CPU operator binary where it was built (oracle/_ref/libref_mpt.so, else oracle/mpt.c),
DeformablePSROIPooling through the C oracle (OpenMP), followed by an SGD-momentum update of every trainable tensor.
Follows symbols/faster/resnet_mx_101_e2e.py:227-345.  kind = "port".
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def make_params(seed=5):
    """Random parameters under the reference's names and shapes (He-normal backbone, N(0,0.01) heads,
    init_weight_rcnn resnet_mx_101_e2e.py:450-485)."""
    import torch_graph as TG
    rng = np.random.RandomState(seed)
    arg, aux = {}, {}

    def conv(n, o, i, k, bias=False, std=None):
        std = (2.0 / (i * k * k)) ** 0.5 if std is None else std
        arg[n + "_weight"] = (rng.standard_normal((o, i, k, k)) * std).astype(np.float32)
        if bias:
            arg[n + "_bias"] = np.zeros(o, np.float32)

    def bn(n, c, var=1.0):
        arg[n + "_gamma"] = np.ones(c, np.float32)
        arg[n + "_beta"] = np.zeros(c, np.float32)
        aux[n + "_moving_mean"] = np.zeros(c, np.float32)
        aux[n + "_moving_var"] = np.full(c, var, np.float32)

    bn("bn_data", 3, 3600.0)
    conv("conv0", 64, 3, 7)
    bn("bn0", 64)
    cin = 64
    for si, n in enumerate(TG.UNITS):
        cout = TG.FILTERS[si + 1]
        for j in range(n):
            nm = "stage%d_unit%d" % (si + 1, j + 1)
            ci, mid = (cin if j == 0 else cout), cout // 4
            bn(nm + "_bn1", ci); conv(nm + "_conv1", mid, ci, 1)
            bn(nm + "_bn2", mid); conv(nm + "_conv2", mid, mid, 3)
            bn(nm + "_bn3", mid); conv(nm + "_conv3", cout, mid, 1)
            if j == 0:
                conv(nm + "_sc", cout, ci, 1)
            if si == 3:
                conv(nm + "_offset", 72, mid, 3, bias=True, std=0.0)
        cin = cout
    conv("rpn_conv_3x3", 512, 3072, 3, True, 0.01); conv("rpn_cls_score", 42, 512, 1, True, 0.01)
    conv("rpn_bbox_pred", 84, 512, 1, True, 0.01); conv("conv_new_1", 256, 3072, 1, True, 0.01)
    for n_, o, i, std in (("offset", 98, 12544, 0.0), ("fc_new_1", 1024, 12544, 0.01), ("fc_new_2", 1024, 1024, 0.01),
                          ("cls_score", 81, 1024, 0.01), ("bbox_pred", 4, 1024, 0.01)):
        arg[n_ + "_weight"] = (rng.standard_normal((o, i)) * std).astype(np.float32)
        arg[n_ + "_bias"] = np.zeros(o, np.float32)
    return arg, aux

def usable_cores():
    """Host threads this process may really use: affinity mask, capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except Exception:
        pass
    return n


def pick_threads(torch, F):
    """All usable cores, unless a short calibration (one 3x3 conv fwd+bwd of the backbone's size) shows that fewer
    threads are faster on this host (oversubscribed / SMT-heavy boxes: 128 threads ran 5x slower than 32)."""
    avail = usable_cores()
    cands = sorted(set([c for c in (8, 16, 32, 64) if c < avail] + [avail]))
    x = torch.randn(1, 256, 64, 64, requires_grad=True)
    w = torch.randn(256, 256, 3, 3, requires_grad=True)
    best, best_t = avail, None
    for c in cands:
        torch.set_num_threads(c)
        ts = []
        for _ in range(3):
            t0 = time.time()
            F.conv2d(x, w, padding=1).sum().backward()
            ts.append(time.time() - t0)
        t = min(ts[1:])
        if best_t is None or t < best_t * 0.9:     # prefer more threads unless clearly slower
            if best_t is None or t < best_t:
                best, best_t = c, t
    return best


def run(sample_chips=1, threads=None, steps=2, warmup=1, budget_s=None):
    """`warmup` untimed + `steps` timed training steps of `sample_chips` chips each (warm oneDNN primitives, warm
    allocator).  budget_s: stop early (never before 2 timed steps) once the timed steps have used that many seconds;
    the returned dict reports the steps actually timed."""
    import torch
    import torch.nn.functional as F
    import oracle_lib as O
    import torch_graph as TG
    from sniper_b200 import synth_batch
    threads = threads or pick_threads(torch, F)
    torch.set_num_threads(threads)
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    B = sample_chips
    arg, aux = make_params()
    P, A = TG.params_to_torch(arg, aux, torch.float32, "cpu")
    mom = {k: torch.zeros_like(v) for k, v in P.items() if v.requires_grad}
    batches = [synth_batch.make_batch(B, seed=100 + i, device="cpu") for i in range(2)]

    def proposals(b):
        def fn(prob, bbox):
            a = (prob.numpy(), bbox.numpy(), b["im_info"].numpy(), b["gt_boxes"].numpy(), b["valid_ranges"].numpy())
            res = O.ref_multi_proposal_target(*a)
            return res if res is not None else O.multi_proposal_target(*a)
        return fn

    def one_step(i, lr=0.0005):
        b = batches[i % len(batches)]
        for v in P.values():
            v.grad = None
        obj, out = TG.forward_train(P, A, b, proposals(b), batch_images=B)
        obj.backward()
        with torch.no_grad():      # SGDMomKernel, optimizer_op-inl.h:279-300
            for k, m in mom.items():
                g = P[k].grad
                if g is None:
                    continue
                wd = 1e-4 if (k.endswith("_weight") or k.endswith("_gamma")) else 0.0
                m.mul_(0.9).add_(P[k], alpha=-lr * wd).add_(g, alpha=-lr)
                P[k].add_(m)
        return float(obj.detach())

    for i in range(warmup):
        one_step(i)
    times = []
    t_all = time.time()
    for i in range(steps):
        t0 = time.time()
        one_step(warmup + i)
        times.append(time.time() - t0)
        if budget_s is not None and len(times) >= 2 and time.time() - t_all > budget_s:
            break
    sec = sum(times)
    n = len(times)
    cpu_model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                cpu_model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return {"value": round(B * n / sec, 4), "unit": "chips/s", "cores": threads, "kind": "port",
            "steps_timed": n, "warmup_steps": warmup, "sec_per_step": round(sec / n, 3),
            "sample": "%d timed step(s) of %d chip(s) of the same workload after %d warm-up step(s), %.1f s "
                      "(PyTorch-CPU fp32 dense layers + bilinear-gather deformable conv + reference CPU "
                      "MultiProposalTarget binary + C-oracle PSROI (OpenMP) + SGD; stand-in for the MXNet CPU stack, "
                      "which cannot be built offline)" % (n, B, warmup, sec),
            "cpu": cpu_model}


def run_mnv2(sample_chips=1, threads=None, steps=2, warmup=1, budget_s=None):
    """The same for BASELINE config 4: MobileNetV2 SNIPER training steps (symbols/faster/mobilenetv2_e2e.py:171-305 as
    restated by oracle/torch_graph_mnv2.py) in float32 on the host cores -- oneDNN dense / depthwise layers, the
    C oracle's MultiProposalTarget at stride 32 / 15 anchors (oracle/mpt.c), C-oracle PSROI (OpenMP), SGD-momentum on every convolution / FC tensor (gamma / beta are FIXED_PARAMS)."""
    import torch
    import torch.nn.functional as F
    import oracle_lib as O
    import torch_graph as TG
    import torch_graph_mnv2 as TM
    from sniper_b200 import synth_batch
    threads = threads or pick_threads(torch, F)
    torch.set_num_threads(threads)
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    B = sample_chips
    arg, aux = TM.make_params(seed=5)
    P, A = TM.params_to_torch(arg, aux, torch.float32, "cpu")
    mom = {k: torch.zeros_like(v) for k, v in P.items() if v.requires_grad}
    batches = [synth_batch.make_batch(B, seed=300 + i, device="cpu", A=15, stride=32) for i in range(2)]
    kw = dict(feat_stride=32, scales=(1, 2, 4, 8, 12), ratios=(0.5, 1, 2))
    TG.MODE[0] = "exact"

    def proposals(b):
        def fn(prob, bbox):
            # (the C oracle, not the reference's CPU operator binary: that one is written for the 21-anchor 32x32 geometry
            #  of the ResNet configuration -- multi_proposal_target-inl.h:118 -- and crashes on 15 anchors at 16x16)
            return O.multi_proposal_target(prob.numpy(), bbox.numpy(), b["im_info"].numpy(), b["gt_boxes"].numpy(),
                                           b["valid_ranges"].numpy(), **kw)
        return fn

    def one_step(i, lr=0.0005):
        b = batches[i % len(batches)]
        for v in P.values():
            v.grad = None
        obj, _ = TM.forward_train(P, A, b, proposals(b), batch_images=B)
        obj.backward()
        with torch.no_grad():
            for k, m in mom.items():
                g = P[k].grad
                if g is None:
                    continue
                wd = 1e-4 if k.endswith("_weight") else 0.0
                m.mul_(0.9).add_(P[k], alpha=-lr * wd).add_(g, alpha=-lr)
                P[k].add_(m)
        return float(obj.detach())

    for i in range(warmup):
        one_step(i)
    times, t_all = [], time.time()
    for i in range(steps):
        t0 = time.time()
        one_step(warmup + i)
        times.append(time.time() - t0)
        if budget_s is not None and len(times) >= 2 and time.time() - t_all > budget_s:
            break
    sec, n = sum(times), len(times)
    return {"value": round(B * n / sec, 4), "unit": "chips/s", "cores": threads, "kind": "port", "steps_timed": n,
            "warmup_steps": warmup, "sec_per_step": round(sec / n, 3),
            "sample": "%d timed step(s) of %d chip(s) of the MobileNetV2 SNIPER step after %d warm-up step(s), %.1f s "
                      "(PyTorch-CPU fp32 dense + depthwise layers, C-oracle MultiProposalTarget and PSROI (OpenMP), SGD; stand-in "
                      "for the MXNet CPU stack)" % (n, B, warmup, sec)}


if __name__ == "__main__":
    import json
    if len(sys.argv) > 1 and sys.argv[1] == "mnv2":
        print(json.dumps(run_mnv2(int(sys.argv[2]) if len(sys.argv) > 2 else 1)))
    else:
        print(json.dumps(run(int(sys.argv[1]) if len(sys.argv) > 1 else 1)))
