"""Layer-by-layer comparison of the product forward with oracle/torch_graph (debug aid for test_graph_parity_gpu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import test_graph_parity_gpu as T
import torch_graph as TG
from sniper_b200 import ops

cfg, net, batch = T._build(2, seed=5)
arg, aux = net.export_reference()
P, A = TG.params_to_torch(arg, aux, torch.float64, "cuda")
taps = {}
TG.MODE[0] = sys.argv[1] if len(sys.argv) > 1 else "tf32"
with torch.no_grad():
    TG.backbone(P, A, batch["data"].double(), taps=taps)
rel = lambda a, b: ((a.double() - b).norm() / b.norm()).item()
x = ops.stem_conv(batch["data"], net.conv0_w, net.bn_data.st.scale, net.bn_data.st.shift, net.bn0.st.scale, net.bn0.st.shift)
print("relu0", rel(x.permute(0, 3, 1, 2), taps["relu0"]))
x = ops.maxpool3x3s2(x)
print("pool0", rel(x.permute(0, 3, 1, 2), taps["pool0"]))
has = False
for i, u in enumerate(net.units):
    nxt = net.units[i + 1].bn1 if i + 1 < len(net.units) else None
    x = u.fwd(x, cfg, x_has_stats=has, next_bn=nxt)
    has = nxt is not None and not nxt.frozen
    t = taps[u.name]
    if u.saved is not None:
        xx, a1, c1, a2, c2, a3, off, col = u.saved
        print(u.name, "a1 %.2e c1 %.2e c2 %.2e out %.2e" % (rel(a1.permute(0, 3, 1, 2), t["a1"]), rel(c1.permute(0, 3, 1, 2), t["c1"]),
              rel(c2.permute(0, 3, 1, 2), t["c2"]), rel(x.permute(0, 3, 1, 2), t["out"])))
    else:
        print(u.name, "out %.2e" % rel(x.permute(0, 3, 1, 2), t["out"]))
