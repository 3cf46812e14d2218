"""ORACLE / CPU baseline -- test & measurement infrastructure only (never on the product path).

Times ONE SNIPER training step (ResNet-101 Faster-R-CNN/R-FCN, 512x512 chips, fp32) on the host cores:
the reference's MXNet CPU stack cannot be built offline, so the dense layers run as the same
graph in PyTorch-CPU fp32 (oneDNN; stand-in for MXNet's im2col+OpenBLAS), the deformable 3x3 as a dilated 3x3
of identical FLOPs plus its offset conv, and the SNIPER operators through the C oracle (oracle/mpt.c,
oracle/psroi.c -- restatements of multi_proposal_target.cu / deformable_psroi_pooling.cu).
Follows symbols/faster/resnet_mx_101_e2e.py:227-345.  kind = "port".
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _build(torch, F):
    g = torch.Generator().manual_seed(5)
    P = {}

    def conv_w(name, cout, cin, k):
        P[name] = (torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5).requires_grad_(True)

    def bn_p(name, c):
        P[name + "_g"] = torch.ones(c, requires_grad=True)
        P[name + "_b"] = torch.zeros(c, requires_grad=True)

    conv_w("conv0", 64, 3, 7)
    units = []
    fl = [64, 256, 512, 1024, 2048]
    cin = 64
    for si, n in enumerate((3, 4, 23, 3)):
        cout = fl[si + 1]
        for j in range(n):
            name = "s%du%d" % (si + 1, j + 1)
            ci = cin if j == 0 else cout
            mid = cout // 4
            stride = 2 if (j == 0 and si in (1, 2)) else 1
            bn_p(name + "bn1", ci); conv_w(name + "c1", mid, ci, 1)
            bn_p(name + "bn2", mid); conv_w(name + "c2", mid, mid, 3)
            bn_p(name + "bn3", mid); conv_w(name + "c3", cout, mid, 1)
            if j == 0:
                conv_w(name + "sc", cout, ci, 1)
            if si == 3:
                conv_w(name + "off", 72, mid, 3)
            units.append((name, si + 1, stride, j > 0))
        cin = cout
    conv_w("rpn", 512, 3072, 3); conv_w("rpn_cls", 42, 512, 1); conv_w("rpn_bbox", 84, 512, 1); conv_w("new1", 256, 3072, 1)
    for n_, o, i in (("off_fc", 98, 12544), ("fc1", 1024, 12544), ("fc2", 1024, 1024), ("cls", 81, 1024), ("bbox", 4, 1024)):
        P[n_] = (torch.randn(o, i, generator=g) * 0.01).requires_grad_(True)
    return P, units


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


def run(sample_chips=1, threads=None):
    import torch
    import torch.nn.functional as F
    import oracle_lib as O
    from sniper_b200 import synth
    threads = threads or pick_threads(torch, F)
    torch.set_num_threads(threads)
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    B = sample_chips
    P, units = _build(torch, F)
    rng = np.random.RandomState(0)
    data = torch.from_numpy((rng.randn(B, 3, 512, 512) * 60).astype(np.float32))
    im_info, vr = synth.chip_meta(B)
    gts = synth.gt_boxes(rng, B)
    label = torch.randint(-1, 2, (B, 21 * 32 * 32))
    t_start = time.time()

    def bn(x, name, train):
        if train:
            return F.relu(F.batch_norm(x, None, None, P[name + "_g"], P[name + "_b"], True, 0.005, 2e-5))
        return F.relu(x * P[name + "_g"].detach().view(1, -1, 1, 1) + P[name + "_b"].detach().view(1, -1, 1, 1))

    x = F.conv2d(data / 60.0, P["conv0"].detach(), stride=2, padding=3)
    x = F.max_pool2d(F.relu(x), 3, 2, 1)
    c4 = None
    for name, stage, stride, match in units:
        train = stage > 1
        ctx = torch.enable_grad() if train else torch.no_grad()
        with ctx:
            a1 = bn(x, name + "bn1", train)
            c1 = F.conv2d(a1, P[name + "c1"])
            a2 = bn(c1, name + "bn2", train)
            if stage == 4:
                _off = F.conv2d(a2, P[name + "off"], padding=2, dilation=2)
                c2 = F.conv2d(a2, P[name + "c2"], padding=2, dilation=2) + 0.0 * _off.mean()
            else:
                c2 = F.conv2d(a2, P[name + "c2"], stride=stride, padding=1)
            a3 = bn(c2, name + "bn3", train)
            sc = x if match else F.conv2d(a1, P[name + "sc"], stride=stride)
            x = F.conv2d(a3, P[name + "c3"]) + sc
        if name == "s3u23":
            c4 = x
    cat = torch.cat([c4, x], 1)
    rpn = F.relu(F.conv2d(cat, P["rpn"], padding=1))
    cls_score = F.conv2d(rpn, P["rpn_cls"])
    bbox_pred = F.conv2d(rpn, P["rpn_bbox"])
    feat = F.relu(F.conv2d(cat, P["new1"]))
    sc2 = cls_score.view(B, 2, -1)
    rpn_loss = F.cross_entropy(sc2, label.clamp(min=-1), ignore_index=-1)
    prob = torch.softmax(sc2, 1).view(B, 42, 32, 32)
    # the reference's OWN CPU operator binary where it was built (oracle/_ref/libref_mpt.so), else the C restatement
    res = O.ref_multi_proposal_target(prob.detach().numpy(), bbox_pred.detach().numpy(), im_info, gts, vr)
    if res is None:
        res = O.multi_proposal_target(prob.detach().numpy(), bbox_pred.detach().numpy(), im_info, gts, vr)
    rois = res["rois"]
    featn = feat.detach().numpy()
    kw = dict(spatial_scale=0.0625, output_dim=256, group_size=1, pooled=7, part_size=7, spp=4, trans_std=0.1)
    off_t, cnt0, _ = O.deform_psroi_fwd(featn, rois, None, no_trans=True, **kw)
    off_t_t = torch.from_numpy(off_t).requires_grad_(True)
    trans = (off_t_t.view(rois.shape[0], -1) @ P["off_fc"].t()).view(-1, 2, 7, 7)
    pooled, cnt1, _ = O.deform_psroi_fwd(featn, rois, trans.detach().numpy(), no_trans=False, **kw)
    pooled_t = torch.from_numpy(pooled).requires_grad_(True)
    fc1 = F.relu(pooled_t.view(rois.shape[0], -1) @ P["fc1"].t())
    fc2 = F.relu(fc1 @ P["fc2"].t())
    cls = fc2 @ P["cls"].t()
    box = fc2 @ P["bbox"].t()
    lab = torch.from_numpy(res["label"]).long()
    loss_head = F.cross_entropy(cls, lab) + (torch.from_numpy(res["bbox_weight"]) * F.smooth_l1_loss(
        box, torch.from_numpy(res["bbox_target"]), reduction="none")).sum() / 3008.0
    loss_head.backward()
    dfeat, dtrans = O.deform_psroi_bwd(pooled_t.grad.numpy(), cnt1, featn, rois, trans.detach().numpy(), no_trans=False, **kw)
    trans.backward(torch.from_numpy(dtrans.astype(np.float32)))
    dfeat2, _ = O.deform_psroi_bwd(off_t_t.grad.numpy(), cnt0, featn, rois, None, no_trans=True, **kw)
    total_dfeat = torch.from_numpy((dfeat + dfeat2).astype(np.float32))
    rpn_l1 = (F.smooth_l1_loss(bbox_pred, torch.zeros_like(bbox_pred), reduction="none") * (torch.rand_like(bbox_pred) > 0.99)).sum() / 256
    (rpn_loss + rpn_l1 + (feat * total_dfeat).sum()).backward()
    with torch.no_grad():      # SGD momentum update on every trainable tensor
        for k, v in P.items():
            if v.grad is not None:
                v.add_(v.grad, alpha=-0.001)
    sec = time.time() - t_start
    cpu_model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                cpu_model = line.split(":", 1)[1].strip()
                break
    except Exception:
        pass
    return {"value": round(B / sec, 4), "unit": "chips/s", "cores": threads, "kind": "port",
            "sample": "%d chip(s) of the same workload, one full training step in %.1f s (PyTorch-CPU fp32 dense layers + "
                      "reference CPU MultiProposalTarget binary + C-oracle PSROI; stand-in for the MXNet CPU stack, which cannot be built offline)" % (B, sec),
            "cpu": cpu_model}


if __name__ == "__main__":
    import json
    print(json.dumps(run(int(sys.argv[1]) if len(sys.argv) > 1 else 1)))
