"""Whole-graph parity: one SNIPER training step (forward + backward, B = 2 chips of 512x512, ResNet-101, deformable
offsets active) on the sm_100a path against a float64 restatement of the reference's `get_symbol_rcnn(is_train=True)`
(oracle/torch_graph.py, symbols/faster/resnet_mx_101_e2e.py:227-345) built from the SAME parameters exported under the
reference's names and layouts (`export_reference`).  Catches wiring errors a smoke test cannot: stride placement, the
shortcut reading act1, the c4|c5 concat order, the fused [4A deltas | 2A scores] RPN head, the (ph,pw,c) flattening in
front of the FCs, the fused cls|bbox head, loss normalisations, the gradient cut below stage 2.

Discrete outputs (graded bit-exact): rois and labels of the product path == the C oracle's MultiProposalTarget
(oracle/mpt.c) run on the product path's own RPN outputs.  The float64 graph is then fed those rois, so both sides
pool the same regions.

Tolerances (TF32 tensor-core products have a 10-bit mantissa, ~5e-4 relative per product; the backbone chains ~100
contractions and 99 train-mode BatchNorm+ReLU whose masks flip on near-zero pre-activations):
  activations  rel. Frobenius <= 1e-2;  loss sums rel <= 2e-3 (|.| <= 1e-3 abs for the tiny R-CNN box loss);
  every parameter gradient rel. Frobenius <= 6e-2, median over the ~330 tensors <= 1.5e-2.
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _build(B, seed, bf16=False):
    import torch
    from sniper_b200 import model, ops, synth_batch
    cfg = model.Cfg()
    cfg.batch_images = B
    if bf16:
        cfg.bf16 = True
    net = model.SniperResNet101(cfg, deform_offset_std=0.01, seed=seed)
    g = torch.Generator(device="cuda")
    g.manual_seed(seed + 1)
    for bn in net._named_bns():          # non-trivial affine parameters / moving statistics everywhere
        if bn.name == "bn_data":
            continue
        bn.st.gamma.copy_(torch.empty(bn.C, device="cuda").uniform_(0.8, 1.2, generator=g))
        bn.st.beta.copy_(torch.empty(bn.C, device="cuda").normal_(0, 0.1, generator=g))
        if bn.frozen:
            bn.st.moving_mean.copy_(torch.empty(bn.C, device="cuda").normal_(0, 0.1, generator=g))
            bn.st.moving_var.copy_(torch.empty(bn.C, device="cuda").uniform_(0.6, 1.6, generator=g))
            ops.bn_frozen(bn.st, cfg.bn_eps)
    batch = synth_batch.make_batch(B, seed=7, device="cuda")
    return cfg, net, batch


def _reference(net, cfg, batch, prob_nchw, bbox_nchw):
    """float64 graph on the exported parameters; proposals = the C oracle on the product path's RPN outputs."""
    import torch
    import oracle_lib as O
    import torch_graph as TG
    arg, aux = net.export_reference()
    res = O.multi_proposal_target(prob_nchw.cpu().numpy(), bbox_nchw.cpu().numpy(), batch["im_info"].cpu().numpy(),
                                  batch["gt_boxes"].cpu().numpy(), batch["valid_ranges"].cpu().numpy())
    P, A = TG.params_to_torch(arg, aux, torch.float64, "cuda")
    b64 = {k: v.double() for k, v in batch.items()}
    taps = {}
    obj, ref = TG.forward_train(P, A, b64, lambda *_: res, batch_images=cfg.batch_images, taps=taps)
    obj.backward()
    return P, ref, res, taps


def test_training_graph_matches_float64_reference():
    import torch
    cfg, net, batch = _build(2, seed=5)
    out = net.forward_backward(batch)
    torch.cuda.synchronize()
    A = cfg.num_anchors
    prob = out["rpn_cls_prob"].permute(0, 3, 1, 2).contiguous()                    # NHWC [..,2A] -> [B,2A,H,W]
    bbox = out["rpn_head"][..., :4 * A].permute(0, 3, 1, 2).contiguous()
    P, ref, res, _ = _reference(net, cfg, batch, prob, bbox)

    # ---- discrete outputs: bit-exact against the oracle on the same RPN outputs
    assert out["rois"].cpu().numpy().tobytes() == res["rois"].tobytes()
    assert np.array_equal(out["label"].cpu().numpy(), res["label"].reshape(-1))
    assert int((res["label"] > 0).sum()) > 0, "test batch yields no foreground roi"

    # ---- activations
    assert _rel(out["cat"].permute(0, 3, 1, 2), ref["relu1"]) < 1e-2
    assert _rel(prob, ref["rpn_cls_prob"]) < 1e-2
    assert _rel(bbox, ref["rpn_bbox_pred"]) < 1e-2
    assert _rel(out["cls_prob"], ref["cls_prob"]) < 1e-2
    # ---- losses (un-normalised sums, as the product path reports them)
    ls, lr = out["losses"][:4].double().cpu(), ref["loss_sums"].cpu()
    print("losses ours", ls.tolist(), "ref", lr.tolist())
    for i in range(4):
        assert abs(ls[i] - lr[i]) <= 2e-3 * abs(lr[i]) + 1e-3, (i, ls[i].item(), lr[i].item())

    # ---- every parameter gradient, in the reference's names and layouts
    garg, _ = net.export_reference(grads=True)
    rows = []
    for name, p in P.items():
        if not p.requires_grad:
            assert name not in garg or "stage1" in name or "bn0" in name or "conv0" in name
            continue
        assert p.grad is not None, name
        ours = torch.from_numpy(garg[name]).cuda()
        assert ours.shape == p.grad.shape, (name, ours.shape, p.grad.shape)
        rows.append((_rel(ours, p.grad), name, p.grad.norm().item()))
    rows.sort(reverse=True)
    print("worst gradient errors:", [(round(r, 4), n) for r, n, _ in rows[:8]])
    med = rows[len(rows) // 2][0]
    print("median gradient error %.4f over %d tensors" % (med, len(rows)))
    assert len(rows) > 300
    for r, name, nrm in rows:
        assert r < 6e-2 or nrm < 1e-9, (name, r, nrm)
    assert med < 1.5e-2
