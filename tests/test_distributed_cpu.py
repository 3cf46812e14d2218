"""N>1 host-side logic on CPU with gloo (world_size 2): the single flat-bucket gradient all-reduce that
replaces the reference's per-key kvstore push/pull, and the reference arm's rank-0-only contract."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from sniper_b200 import model
    P = model.ParamStore()
    P.bucket = 1
    P.add("s_weight", (6, 2)); P.add("s_beta", (6,))
    P.bucket = 0
    P.add("a_weight", (5, 3)); P.add("a_bias", (5,)); P.add("offset_weight", (7,), lr_mult=0.01); P.add("b_gamma", (4,))
    P.finalize("cpu")
    # every rank owns different chips -> different gradients; the update must see their SUM (rescale_grad=1)
    for i, (name, g) in enumerate(sorted(P.grads.items())):
        g.fill_(float(rank + 1) * (i + 1))
    # one asynchronous collective per gradient bucket, in the order the backward pass completes them (the trainer
    # starts bucket k while phase k+1 computes); together they cover the whole buffer exactly once
    works = [dist.all_reduce(P.g[a:b], op=dist.ReduceOp.SUM, async_op=True) for a, b in P.bucket_ranges]
    for w in works:
        w.wait()
    cover = sorted(P.bucket_ranges)
    assert cover[0][0] == 0 and cover[-1][1] == P.total and all(x[1] == y[0] for x, y in zip(cover, cover[1:]))
    exp = {name: float(sum(r + 1 for r in range(world)) * (i + 1)) for i, name in enumerate(sorted(P.grads))}
    ok = all(bool((P.grad(n) == v).all()) for n, v in exp.items())
    # optimizer groups: (lr_mult, wd_mult) as MXNet assigns them
    groups = sorted(set(g for _, _, g in P.segments))
    ok = ok and groups == [(0.01, 1.0), (1.0, 0.0), (1.0, 1.0)]
    ok = ok and all(any(a <= s0 and e0 <= b for a, b in P.bucket_ranges) for s0, e0, _ in P.segments)
    ok = ok and P.total % 4 == 0 and all(o % 4 == 0 for o, _ in P.layout.values())
    q.put((rank, ok))
    dist.destroy_process_group()


def test_flat_bucket_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29611, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


def test_reference_arm_prints_once_under_torchrun():
    """bench.py --impl reference under a 2-rank launch: rank 0 runs the CPU port and prints ONE JSON line."""
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = "4"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29612", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "chips/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def _mnv2_worker(rank, world, port, q):
    """Data-parallel MobileNetV2 step (config 4) on two gloo ranks: same weights, different chips, ONE all-reduce over the
    model's single gradient bucket (sum, rescale_grad = 1), identical fused SGD update on every rank."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    torch.set_default_dtype(torch.float64)
    torch.set_num_threads(2)
    import fake_ops
    from sniper_b200 import model_mnv2 as MM
    from sniper_b200 import ops, synth_batch

    class _MP:
        def setattr(self, m, k, v):
            setattr(m, k, v)
    fake_ops.install(_MP(), ops)
    cfg = MM.MCfg()
    cfg.batch_images, cfg.bf16, cfg.wgrad_stream = 1, False, False

    def grads_of(seed):
        net = MM.SniperMobileNetV2(cfg, device="cpu", seed=3)
        b = synth_batch.make_batch(1, seed=seed, device="cpu", chip=256, A=15, stride=32)
        net.forward_backward({k: v.double() for k, v in b.items()})
        return net
    net = grads_of(50 + rank)                                   # this rank's chip
    assert len(net.P.bucket_ranges) == 1 and tuple(net.P.bucket_ranges[0]) == (0, net.P.total)
    a, b = net.P.bucket_ranges[0]
    dist.all_reduce(net.P.g[a:b], op=dist.ReduceOp.SUM)
    ok = True
    if rank == 0:                                                # == the two chips' gradients summed in one process
        want = grads_of(50).P.g + grads_of(51).P.g
        ok = bool(torch.allclose(net.P.g, want, rtol=1e-12, atol=1e-14))
    net.update(lr=0.01)
    w = net.P.w.clone()
    gathered = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(gathered, w)
    ok = ok and all(torch.equal(gathered[0], t) for t in gathered)       # replicas stay identical
    q.put((rank, ok))
    dist.destroy_process_group()


def test_mobilenet_data_parallel_step_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_mnv2_worker, args=(r, 2, 29613, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]
