"""RPN-only test graph (get_symbol_rpn, is_train=False; resnet_mx_101_e2e.py:157-225) on the GPU.  Kept in the LAST test
file of the suite: it was added after this round's GPU budget was spent (its launches are a prefix of the validated
forward_inference), and `pytest -x` must not be able to stop earlier files on it."""
import pytest

pytestmark = pytest.mark.gpu

def test_rpn_only_inference_equals_the_proposal_half_of_the_full_graph():
    """get_symbol_rpn(is_train=False) as SniperResNet101.forward_rpn: the same launches as forward_inference up to the
    proposal operator, so rois and scores are bit-identical to the full test graph's."""
    import torch
    from sniper_b200 import model, synth_batch
    cfg = model.Cfg()
    cfg.batch_images = 2
    net = model.SniperResNet101(cfg, deform_offset_std=0.01)
    batch = synth_batch.make_batch(2, seed=8, device="cuda")
    net.train_step(batch, lr=0.001)
    rois, scores, _, _ = net.forward_inference(batch["data"], batch["im_info"])
    r2, s2 = net.forward_rpn(batch["data"], batch["im_info"])
    torch.cuda.synchronize()
    assert torch.equal(rois, r2) and torch.equal(scores, s2)
