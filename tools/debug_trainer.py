import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sniper_b200 import model, synth_batch, trainer
batch = synth_batch.make_batch(1, seed=4, device="cpu", pinned=True)
trs = []
for use_graph in (True, False, False):
    cfg = model.Cfg(); cfg.batch_images = 1
    trs.append(trainer.Trainer(cfg, use_graph=use_graph, seed=5, deform_offset_std=0.01))
w0 = trs[0].net.P.w.clone()
print("init equal", [torch.equal(w0, t.net.P.w) for t in trs])
for lr in (0.001, 0.004, 0.0005):
    ls = [t.step(batch, lr=lr) for t in trs]
    torch.cuda.synchronize()
    g = [t.net.P.g.double() for t in trs]
    w = [(t.net.P.w - w0).double() for t in trs]
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()
    print("lr", lr, "grad graph-vs-eager %.3e eager-vs-eager %.3e | dw %.3e %.3e | gnorm %.3e" % (rel(g[0], g[1]), rel(g[2], g[1]), rel(w[0], w[1]), rel(w[2], w[1]), g[1].norm().item()))
    print("   losses", [round(l["rcnn_cls_loss"], 3) for l in ls], [round(l["rpn_cls_loss"], 3) for l in ls])
