"""Whole-graph parity: one SNIPER training step (forward + backward, B = 2 chips of 512x512, ResNet-101, deformable
offsets active) on the sm_100a path against a float64 restatement of the reference's `get_symbol_rcnn(is_train=True)`
(oracle/torch_graph.py, symbols/faster/resnet_mx_101_e2e.py:227-345) built from the SAME parameters exported under the
reference's names and layouts (`export_reference`).  Catches wiring errors a smoke test cannot: stride placement, the
shortcut reading act1, the c4|c5 concat order, the fused [4A deltas | 2A scores] RPN head, the (ph,pw,c) flattening in
front of the FCs, the fused cls|bbox head, loss normalisations, the gradient cut below stage 2.

Discrete outputs (graded bit-exact): rois and labels of the product path == the C oracle's MultiProposalTarget
(oracle/mpt.c) run on the product path's own RPN outputs.  The float64 graph is then fed those rois, so both sides
pool the same regions.

Two float64 references (oracle/torch_graph.MODE):
  "tf32"  -- the operands of every tensor-core contraction (forward, data and weight gradients) are reduced to TF32 by
             dropping the low 13 mantissa bits, exactly what `tcgen05.mma kind::tf32` does with fp32 words
             (test_tcgen05_tf32_reads_truncate measures it), products accumulated exactly.  This is the sharp test: the
             product path must agree to accumulation-order noise.  Tolerances: activations 2e-3 rel. Frobenius, loss sums
             1e-3 rel, EVERY parameter gradient 2e-2 rel. Frobenius (median 3e-3).
  "exact" -- real arithmetic.  Documents the cost of TF32 itself: a random-init ResNet-101 amplifies each residual
             unit's ~1.5e-3 TF32 error by a few percent per unit (measured: 4e-3 after stage 1, 9e-2 at c4, 3e-1 at c5),
             so only loose bounds make sense here: losses 5e-2 rel, RPN probabilities 5e-2.
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
        # bn3 scales the residual branch (a3 = relu(bn3(.)) feeds conv3, whose output is added to the shortcut): a
        # small gamma makes every unit a near-identity map, as in a trained network.  With gamma ~ 1 a He-initialised
        # ResNet-101 is chaotic -- it amplifies ANY perturbation by ~15 % per residual unit (measured: two float64
        # evaluations that differ only by TF32 operand truncation are 9 % apart at c4, 30 % at c5, and their parameter
        # gradients are uncorrelated), so no tolerance could separate a wiring error from rounding.
        lo, hi, sd = (0.15, 0.25, 0.02) if bn.name.endswith("_bn3") else (0.8, 1.2, 0.1)
        bn.st.gamma.copy_(torch.empty(bn.C, device="cuda").uniform_(lo, hi, generator=g))
        bn.st.beta.copy_(torch.empty(bn.C, device="cuda").normal_(0, sd, generator=g))
        if bn.frozen:
            bn.st.moving_mean.copy_(torch.empty(bn.C, device="cuda").normal_(0, 0.1, generator=g))
            bn.st.moving_var.copy_(torch.empty(bn.C, device="cuda").uniform_(0.6, 1.6, generator=g))
            ops.bn_frozen(bn.st, cfg.bn_eps)
    batch = synth_batch.make_batch(B, seed=7, device="cuda")
    return cfg, net, batch


# Gradient tolerances are set by ReLU: a forward deviation eps (relative to the activation scale) flips the sign of a
# fraction ~0.4*eps of the pre-activations, and every flipped unit changes its gradient by 100 %, so gradients agree
# only to ~sqrt(0.4*eps) per ReLU layer (eps = 4e-3 -> 4 %), accumulating over the ~100 ReLUs below the heads.  A wiring
# error gives relative errors >= 0.7 (uncorrelated gradients) and norm ratios far from 1.
TOL_TF32 = dict(act=1e-2, loss=5e-3, grad=0.25, grad_median=0.15, head=8e-2, norm=0.05)
# bf16: every stored activation is rounded to 8 bits (2e-3); rounding is discontinuous, so product and emulation part
# ways at that level per tensor (measured: 3e-2 at c4|c5) and the ReLU argument above gives ~0.35 on backbone gradients
TOL_BF16 = dict(act=5e-2, loss=3e-2, grad=0.6, grad_median=0.45, head=0.25, norm=0.1)


def _tf32_trunc(x):
    import torch
    return (x.float().contiguous().view(torch.int32) & -8192).view(torch.float32).double()


def _tf32_rne(x):
    import torch
    i = x.float().contiguous().view(torch.int32)
    i = i + 0x0FFF + ((i >> 13) & 1)
    return (i & -8192).view(torch.float32).double()


def test_tcgen05_tf32_reads_truncate():
    """kind::tf32 ignores the low 13 mantissa bits of the fp32 words it reads (no rounding): the product GEMM equals
    the exact product of TRUNCATED operands to accumulation noise, and is ~100x further from round-to-nearest operands."""
    import torch
    from sniper_b200 import ops
    torch.manual_seed(0)
    a = torch.randn(512, 1024, device="cuda")
    b = torch.randn(256, 1024, device="cuda")
    y = ops.gemm_nt(a, b).double()
    yt = _tf32_trunc(a) @ _tf32_trunc(b).t()
    yr = _tf32_rne(a) @ _tf32_rne(b).t()
    et, er = _rel(y, yt), _rel(y, yr)
    print("vs truncated operands %.2e, vs RNE operands %.2e" % (et, er))
    assert et < 5e-6 and er > 20 * et


def _reference(net, cfg, batch, prob_nchw, bbox_nchw, mode):
    """float64 graph on the exported parameters; proposals = the C oracle on the product path's RPN outputs."""
    import torch
    import oracle_lib as O
    import torch_graph as TG
    arg, aux = net.export_reference()
    res = O.multi_proposal_target(prob_nchw.cpu().numpy(), bbox_nchw.cpu().numpy(), batch["im_info"].cpu().numpy(),
                                  batch["gt_boxes"].cpu().numpy(), batch["valid_ranges"].cpu().numpy())
    P, A = TG.params_to_torch(arg, aux, torch.float64, "cuda")
    b64 = {k: v.double() for k, v in batch.items()}
    TG.MODE[0] = mode
    try:
        obj, ref = TG.forward_train(P, A, b64, lambda *_: res, batch_images=cfg.batch_images)
        obj.backward()
    finally:
        TG.MODE[0] = "exact"
    return P, ref, res


@pytest.mark.parametrize("bf16", [False, True])
def test_training_graph_matches_float64_reference(bf16):
    import torch
    cfg, net, batch = _build(2, seed=5, bf16=bf16)
    out = net.forward_backward(batch)
    torch.cuda.synchronize()
    A = cfg.num_anchors
    prob = out["rpn_cls_prob"].permute(0, 3, 1, 2).contiguous()                    # NHWC [..,2A] -> [B,2A,H,W]
    bbox = out["rpn_head"][..., :4 * A].permute(0, 3, 1, 2).contiguous()
    garg, _ = net.export_reference(grads=True)
    ls = out["losses"][:4].double().cpu()

    # ================= sharp: float64 with TF32-truncated contraction operands
    P, ref, res = _reference(net, cfg, batch, prob, bbox, "bf16" if bf16 else "tf32")
    # ---- discrete outputs: bit-exact against the oracle on the same RPN outputs
    assert out["rois"].cpu().numpy().tobytes() == res["rois"].tobytes()
    assert np.array_equal(out["label"].cpu().numpy(), res["label"].reshape(-1))
    assert int((res["label"] > 0).sum()) > 0, "test batch yields no foreground roi"
    acts = dict(cat=_rel(out["cat"].permute(0, 3, 1, 2), ref["relu1"]), rpn_prob=_rel(prob, ref["rpn_cls_prob"]),
                rpn_bbox=_rel(bbox, ref["rpn_bbox_pred"]), cls_prob=_rel(out["cls_prob"], ref["cls_prob"]))
    print("activation errors (%s reference):" % ("bf16" if bf16 else "tf32"), {k: "%.2e" % v for k, v in acts.items()})
    lr = ref["loss_sums"].cpu()
    print("losses ours", ls.tolist(), "tf32-ref", lr.tolist())
    rows = []
    for name, p in P.items():
        if not p.requires_grad:
            assert name not in garg
            continue
        assert p.grad is not None, name
        ours = torch.from_numpy(garg[name]).cuda()
        assert ours.shape == p.grad.shape, (name, ours.shape, p.grad.shape)
        rows.append((_rel(ours, p.grad), name, p.grad.norm().item(), ours.double().norm().item()))
    if os.environ.get("SNIPER_GRAD_TABLE"):
        with open(os.environ["SNIPER_GRAD_TABLE"], "w") as f:
            for r in rows:
                f.write("%-40s rel %.3e  |ref| %.3e  |ours| %.3e\n" % (r[1], r[0], r[2], r[3]))
    rows.sort(reverse=True)
    print("worst gradient errors:", [(round(r[0], 5), r[1]) for r in rows[:8]])
    med = rows[len(rows) // 2][0]
    print("median gradient error %.2e over %d tensors" % (med, len(rows)))
    assert len(rows) == 297
    tol = TOL_BF16 if getattr(cfg, "bf16", False) else TOL_TF32
    for k, v in acts.items():
        assert v < tol["act"], (k, v)
    for i in range(4):
        assert abs(ls[i] - lr[i]) <= tol["loss"] * abs(lr[i]) + 1e-4, (i, ls[i].item(), lr[i].item())
    for r, name, nrm, ours_n in rows:
        assert r < tol["grad"] or nrm < 1e-9, (name, r, nrm)
        assert abs(ours_n / nrm - 1) < tol["norm"] or nrm < 1e-9, (name, ours_n, nrm)
        if "stage" not in name:
            assert r < tol["head"], (name, r)
    assert med < tol["grad_median"]

    # ================= loose: real arithmetic (what TF32 costs on this network)
    P2, ref2, _ = _reference(net, cfg, batch, prob, bbox, "exact")
    l2 = ref2["loss_sums"].cpu()
    e_prob = _rel(prob, ref2["rpn_cls_prob"])
    g2 = sorted(_rel(torch.from_numpy(garg[n]).cuda(), p.grad) for n, p in P2.items() if p.requires_grad)
    print("exact reference: losses", l2.tolist(), "rpn_prob err %.2e, c4|c5 err %.2e, gradient err median %.2e max %.2e"
          % (e_prob, _rel(out["cat"].permute(0, 3, 1, 2), ref2["relu1"]), g2[len(g2) // 2], g2[-1]))
    assert e_prob < 5e-2
    for i in range(4):
        assert abs(ls[i] - l2[i]) <= 5e-2 * abs(l2[i]) + 1e-2, (i, ls[i].item(), l2[i].item())
