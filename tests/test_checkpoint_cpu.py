"""MXNet `.params` wire format and the reference <-> repo layout mapping (CPU only)."""
import os
import struct

import numpy as np
import pytest

from sniper_b200 import checkpoint as C


def test_params_file_format_roundtrip_and_known_bytes(tmp_path):
    rng = np.random.RandomState(0)
    arg = {"conv0_weight": rng.randn(4, 3, 7, 7).astype(np.float32), "half": rng.randn(5).astype(np.float16)}
    aux = {"bn0_moving_mean": rng.randn(4).astype(np.float32)}
    p = str(tmp_path / "model-0007.params")
    C.write_params(p, arg, aux)
    raw = open(p, "rb").read()
    # ndarray.cc:1748-1760: list magic 0x112, reserved 0, count; ndarray.cc:1547-1560: V2 magic, stype 0, shape (u32 ndim + i64 dims)
    assert struct.unpack_from("<QQQ", raw, 0) == (0x112, 0, 3)
    assert struct.unpack_from("<Ii", raw, 24) == (0xF993FAC9, 0)
    assert struct.unpack_from("<I4q", raw, 32) == (4, 4, 3, 7, 7)
    assert struct.unpack_from("<iii", raw, 32 + 4 + 32) == (1, 0, 0)          # cpu(0), float32
    a2, x2 = C.read_params(p)
    assert set(a2) == set(arg) and set(x2) == set(aux)
    for k in arg:
        assert a2[k].dtype == arg[k].dtype and np.array_equal(a2[k], arg[k])
    assert np.array_equal(x2["bn0_moving_mean"], aux["bn0_moving_mean"])


def test_reads_v1_and_legacy_ndarrays(tmp_path):
    """NDArray::LegacyLoad (ndarray.cc:1618-1665): V1 magic + int64 dims, and the magic-less form with uint32 dims."""
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    name = b"arg:w"
    body_v1 = struct.pack("<I", 0xF993FAC8) + struct.pack("<I2q", 2, 2, 3) + struct.pack("<iii", 1, 0, 0) + a.tobytes()
    body_legacy = struct.pack("<I2I", 2, 2, 3) + struct.pack("<iii", 2, 0, 0) + a.tobytes()
    for i, body in enumerate((body_v1, body_legacy)):
        p = str(tmp_path / ("old%d.params" % i))
        with open(p, "wb") as f:
            f.write(struct.pack("<QQQ", 0x112, 0, 1) + body + struct.pack("<QQ", 1, len(name)) + name)
        arg, aux = C.read_params(p)
        assert np.array_equal(arg["w"], a) and not aux


def test_conv_and_fc_layout_conversions_preserve_the_operator():
    """OIHW -> tap-major rows and NCHW-flat -> NHWC-flat FC weights compute the same function (checked with PyTorch CPU
    as the NCHW reference), and the inverse conversions restore the reference tensors exactly."""
    import torch
    import torch.nn.functional as F
    rng = np.random.RandomState(1)
    w = rng.randn(6, 5, 3, 3).astype(np.float32)
    x = rng.randn(2, 5, 8, 8).astype(np.float32)
    rows = C.conv_to_rows(w, rows=8)
    assert rows.shape == (8, 45) and not rows[6:].any()
    ref = F.conv2d(torch.from_numpy(x), torch.from_numpy(w), padding=1).numpy()
    xp = np.pad(x.transpose(0, 2, 3, 1), ((0, 0), (1, 1), (1, 1), (0, 0)))
    patches = np.stack([xp[:, i:i + 8, j:j + 8, :] for i in range(3) for j in range(3)], 3).reshape(2, 8, 8, 45)
    mine = patches @ rows[:6].T                                     # NHWC implicit GEMM, K = (tap, channel)
    assert np.abs(mine.transpose(0, 3, 1, 2) - ref).max() < 1e-4
    assert np.array_equal(C.rows_to_conv(rows, 6, 5, 3), w)
    fc = rng.randn(10, 4 * 7 * 7).astype(np.float32)
    pooled = rng.randn(3, 4, 7, 7).astype(np.float32)
    m = C.fc_chw_to_hwc(fc, 4, 7, 7, rows=16)
    a = pooled.reshape(3, -1) @ fc.T
    b = pooled.transpose(0, 2, 3, 1).reshape(3, -1) @ m[:10].T
    assert np.abs(a - b).max() < 1e-4
    assert np.array_equal(C.fc_hwc_to_chw(m, 10, 4, 7, 7), fc)


def test_fused_heads_and_pooled_fcs_roundtrip():
    rng = np.random.RandomState(2)
    A, K = 21, 81
    arg = {"rpn_bbox_pred_weight": rng.randn(4 * A, 512, 1, 1).astype(np.float32), "rpn_bbox_pred_bias": rng.randn(4 * A).astype(np.float32),
           "rpn_cls_score_weight": rng.randn(2 * A, 512, 1, 1).astype(np.float32), "rpn_cls_score_bias": rng.randn(2 * A).astype(np.float32),
           "cls_score_weight": rng.randn(K, 1024).astype(np.float32), "cls_score_bias": rng.randn(K).astype(np.float32),
           "bbox_pred_weight": rng.randn(4, 1024).astype(np.float32), "bbox_pred_bias": rng.randn(4).astype(np.float32),
           "offset_weight": rng.randn(98, 256 * 49).astype(np.float32), "offset_bias": rng.randn(98).astype(np.float32)}
    w, b = C.conv_from_reference("rpn_head", 6 * A, 128, 512, 1, True, arg)
    assert w.shape == (128, 512) and np.array_equal(w[:4 * A], arg["rpn_bbox_pred_weight"].reshape(4 * A, 512))
    assert np.array_equal(w[4 * A:6 * A], arg["rpn_cls_score_weight"].reshape(2 * A, 512)) and not w[6 * A:].any()
    out = C.conv_to_reference("rpn_head", 6 * A, 512, 1, w, b, (4 * A, 2 * A))
    w2, b2 = C.conv_from_reference("cls_bbox", K + 4, 96, 1024, 1, True, arg)
    C.conv_to_reference("cls_bbox", K + 4, 1024, 1, w2, b2, (K, 4), out)
    w3, b3 = C.conv_from_reference("offset", 98, 128, 256 * 49, 1, True, arg)
    C.conv_to_reference("offset", 98, 256 * 49, 1, w3, b3, None, out)
    for k in arg:
        assert np.array_equal(out[k], arg[k]), k
    with pytest.raises(ValueError):
        C.conv_from_reference("cls_bbox", K + 4, 96, 512, 1, True, arg)


def test_save_checkpoint_adds_test_time_bbox_weights(tmp_path):
    """checkpoint_callback (resnet_mx_101_e2e.py:6-17): `bbox_pred_*_test` = weights / bias scaled by (0.1, 0.1, 0.2, 0.2);
    load_param(process=True) (utils.py:77-100) renames them over the training tensors."""
    import numpy as np
    from sniper_b200 import checkpoint as ck
    rng = np.random.RandomState(0)
    arg = {"bbox_pred_weight": rng.randn(4, 1024).astype(np.float32), "bbox_pred_bias": rng.randn(4).astype(np.float32),
           "cls_score_weight": rng.randn(81, 1024).astype(np.float32)}
    aux = {"bn0_moving_mean": rng.randn(64).astype(np.float32)}
    path = ck.save_checkpoint(str(tmp_path / "m"), 6, arg, aux)
    assert path.endswith("m-0006.params")
    a, x = ck.load_param(str(tmp_path / "m"), 6)
    stds = np.array([0.1, 0.1, 0.2, 0.2], np.float32)
    assert np.array_equal(a["bbox_pred_weight_test"], (arg["bbox_pred_weight"].T * stds).T)
    assert np.array_equal(a["bbox_pred_bias_test"], arg["bbox_pred_bias"] * stds)
    assert np.array_equal(a["bbox_pred_weight"], arg["bbox_pred_weight"]) and np.array_equal(x["bn0_moving_mean"], aux["bn0_moving_mean"])
    a2, _ = ck.load_param(str(tmp_path / "m"), 6, process=True)
    assert "bbox_pred_weight_test" not in a2 and np.array_equal(a2["bbox_pred_weight"], a["bbox_pred_weight_test"])
