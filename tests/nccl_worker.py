"""Worker of tests/test_nccl_gpu.py (one process per GPU, launched by torch.distributed.run)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(out_path):
    from sniper_b200 import model, synth_batch, trainer
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    per = 2
    cfg = model.Cfg()
    cfg.batch_images = per
    tr = trainer.Trainer(cfg, device="cuda:%d" % local, world_size=world, use_graph=False, seed=5, deform_offset_std=0.01)
    full = synth_batch.make_batch(per * world, seed=11, device="cpu")
    mine = {k: v[rank * per:(rank + 1) * per].contiguous().pin_memory() for k, v in full.items()}
    tr.load(mine)
    P = tr.net.P
    # (1) the collective: bucket after the all-reduce == sum of the rank-local buckets, bit for bit
    tr.net.forward_backward(tr.static)
    local_g = P.g.clone()
    tr._allreduce()
    gathered = [torch.empty_like(local_g) for _ in range(world)] if rank == 0 else None
    dist.gather(local_g, gathered, dst=0)
    res = {}
    if rank == 0:
        total = gathered[0].clone()
        for t in gathered[1:]:
            total += t
        res["allreduce_equals_sum_of_locals"] = bool(torch.equal(total, P.g)) if world == 2 else \
            float((total - P.g).abs().max() / P.g.abs().max()) < 1e-6
        # (2) data-parallel semantics: the same chips processed one slice after the other on ONE GPU (own BatchNorm
        # statistics per slice, as per-GPU BN demands) give the same summed gradient up to float-atomic noise
        ref = torch.zeros_like(P.g)
        for r in range(world):
            sl = {k: v[r * per:(r + 1) * per].contiguous().cuda(local) for k, v in full.items()}
            tr.net.forward_backward(sl)
            ref += P.g
        P.g.copy_(total)
        res["vs_sequential_rel"] = float((ref - total).norm() / total.norm())
    # (3) replicated update: every rank ends with bit-identical weights
    tr.net.update(lr=0.001)
    cs = torch.stack([P.w.double().sum(), P.w.double().abs().sum(), P.mom.double().abs().sum()])
    allcs = [torch.empty_like(cs) for _ in range(world)]
    dist.all_gather(allcs, cs)
    if rank == 0:
        res["weights_identical"] = all(torch.equal(allcs[0], c) for c in allcs[1:])
        res["world"] = world
        json.dump(res, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
