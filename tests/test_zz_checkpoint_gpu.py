"""Reference-checkpoint round trip on the device (kept in its own file so that it runs after every other GPU test)."""
import pytest

pytestmark = pytest.mark.gpu


def test_reference_checkpoint_roundtrip(tmp_path):
    """export_reference -> MXNet .params file -> read_params -> load_reference restores every tensor (reference names and
    layouts: OIHW, NCHW-flattened FCs, split heads) and the inference outputs."""
    import numpy as np
    import torch
    from sniper_b200 import checkpoint, model, synth_batch
    cfg = model.Cfg()
    cfg.batch_images = 1
    net = model.SniperResNet101(cfg, deform_offset_std=0.01, seed=5)
    batch = synth_batch.make_batch(1, seed=3, device="cuda")
    net.train_step(batch, lr=0.001)
    arg, aux = net.export_reference()
    assert arg["stage2_unit1_conv2_weight"].shape == (128, 128, 3, 3) and arg["rpn_cls_score_weight"].shape == (42, 512, 1, 1)
    assert arg["fc_new_1_weight"].shape == (1024, 256 * 7 * 7) and arg["bbox_pred_weight"].shape == (4, 1024)
    assert "stage3_unit5_bn2_moving_var" in aux and "bn_data_moving_mean" in aux
    # the file format itself is covered on the CPU (tests/test_checkpoint_cpu.py); here only the heads travel through
    # a file (a full checkpoint is 300 MB), the rest is handed over in memory
    p = str(tmp_path / "sniper-0001.params")
    small = {k: v for k, v in arg.items() if k.startswith(("rpn_", "cls_score", "bbox_pred", "fc_new_2"))}
    checkpoint.write_params(p, small, {})
    ref_out = net.forward_inference(batch["data"], batch["im_info"])
    net2 = model.SniperResNet101(cfg, deform_offset_std=0.0, seed=11)          # different weights
    a2, x2 = checkpoint.read_params(p)
    a2 = dict(arg, **a2)
    x2 = dict(aux)
    net2.load_reference(a2, x2)
    arg3, aux3 = net2.export_reference()
    for k in arg:
        assert np.array_equal(arg[k], arg3[k]), k
    for k in aux:
        assert np.array_equal(aux[k], aux3[k]), k
    out2 = net2.forward_inference(batch["data"], batch["im_info"])
    assert all(torch.equal(a, b) for a, b in zip(ref_out, out2))


def test_backbone_only_checkpoint_loads_with_allow_missing():
    """The reference's normal training start: an ImageNet ResNet-101 `.params` has no RPN / R-FCN / offset tensors
    (init_weight_rcnn fills them, resnet_mx_101_e2e.py:450-485)."""
    import torch
    from sniper_b200 import model
    cfg = model.Cfg()
    net = model.SniperResNet101(cfg, seed=3)
    arg, aux = net.export_reference()
    heads = ("rpn_", "conv_new_1", "offset", "fc_new", "cls_score", "bbox_pred")
    backbone = {k: v for k, v in arg.items() if not any(h in k for h in heads)}
    net2 = model.SniperResNet101(cfg, seed=9)
    keep = net2.fc_new_1.w.clone()
    with pytest.raises(KeyError):
        net2.load_reference(backbone, aux)
    skipped = net2.load_reference(backbone, aux, allow_missing=True)
    assert "rpn_head" in skipped and "cls_bbox" in skipped and "stage4_unit1_offset" in skipped and "fc_new_1" in skipped
    assert not any(s.startswith("stage3") for s in skipped)
    assert torch.equal(net2.fc_new_1.w, keep)                                           # untouched initialisation
    assert torch.equal(net2.units[10].conv1.w, net.units[10].conv1.w)                   # backbone taken from the file
    assert float(net2.units[-1].offset.w.abs().sum()) == 0.0                            # offset layers stay zero
