"""Host-only production rate of MNIteratorE2E (no GPU work): ms per batch of 20 chips, plain and through PrefetchingIter
while the main thread sleeps; plus the per-phase split of one training step fed by it (when a GPU is present)."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sniper_b200 import iterator as IT

cfg = IT.default_config()
roidb = IT.synthetic_roidb(24, seed=11, n_prop=300)
it = IT.MNIteratorE2E(roidb, cfg, batch_size=20, n_buffers=6)
pf = IT.PrefetchingIter(it, depth=3)
for _ in range(12):       # every staging buffer allocated, page-faulted and (with a GPU) pinned
    next(pf)
t = time.time()
n = 30
for _ in range(n):
    next(pf)
print("host only, prefetch thread: %.1f ms/batch, cores %d" % ((time.time() - t) / n * 1e3, os.cpu_count()))
if torch.cuda.is_available():
    from sniper_b200 import model
    from sniper_b200.trainer import Trainer
    from sniper_b200 import synth_batch
    c = model.Cfg(); c.batch_images = 20
    tr = Trainer(c, device="cuda:0", world_size=1, use_graph=True)
    tr.load(synth_batch.make_batch(20, seed=100, device="cpu", pinned=True))
    tr.capture()
    stage = IT.InputStage(cfg, "cuda:0", 20)
    for _ in range(8):
        tr.step_raw(next(pf), stage)
    torch.cuda.synchronize()
    tw = ts = 0.0
    t0 = time.time()
    for _ in range(n):
        a = time.time(); raw = next(pf); b = time.time()
        tr.step_raw(raw, stage); c2 = time.time()
        tw += b - a; ts += c2 - b
    torch.cuda.synchronize()
    print("with training: %.1f ms/step total, %.1f waiting for the iterator, %.1f in step_raw" %
          ((time.time() - t0) / n * 1e3, tw / n * 1e3, ts / n * 1e3))
pf.close()
