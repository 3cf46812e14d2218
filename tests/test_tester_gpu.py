"""AutoFocus inference pyramid on the device: rectangular input canvas, inference forward at non-512 sizes with the
FocusPixel branch, FocusChips feeding the next scale, device vs host soft-NMS aggregation."""
from types import SimpleNamespace

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _small_cfg():
    from sniper_b200 import iterator as IT
    cfg = IT.default_config()
    cfg.TEST.SCALES = [(96, 160), (192, 320), (288, 480)]          # same structure as (480,512) / (800,1280) / (1400,2000)
    cfg.TEST.BATCH_IMAGES = [2, 2, 1]
    cfg.TEST.VALID_RANGES = [(40, -1), (16, 90), (-1, 40)]
    cfg.TEST.CHIP_HYPERPARAMS = [(3, 0.3, 4), (3, 0.4, 5)]
    return cfg


def test_rectangular_canvas_equals_square_kernel_on_its_overlap():
    import torch
    from sniper_b200 import iterator as IT
    from sniper_b200 import tester as TS
    from sniper_b200._lib import check, lib
    cfg = IT.default_config()
    rng = np.random.RandomState(1)
    db = [dict(width=300, height=200, image_data=rng.randint(0, 256, (200, 300, 3)).astype(np.uint8),
               inference_crops=np.array([[0, 0, 300, 200], [50, 20, 250, 180]]))]
    it = TS.MNIteratorTestAutoFocus(db, cfg, (160, 256), batch_size=2)
    b = next(it)
    data = b['data']
    B, _, SH, SW = data.shape
    assert (SH, SW) == (160, 256)                   # scale 0.8: 160 x 240 -> canvas 160 x 256
    S = max(SH, SW)
    # the square entry point with the same table rows: identical pixels where the canvases overlap, zeros elsewhere
    sq = torch.empty(B, 3, S, S, device="cuda")
    tab = torch.zeros(B, 8, dtype=torch.int64)
    off = 0
    for k, c in enumerate(db[0]['inference_crops'][[int(i) for i in b['chip_ids']]]):
        h, w = int(c[3]) - int(c[1]), int(c[2]) - int(c[0])
        s = float(b['scales'][k])
        tab[k] = torch.tensor([off, h, w, int(np.rint(h * s)), int(np.rint(w * s)), 0, int(np.float64(s).view(np.int64)), 0])
        off += h * w * 3
    check(lib().sniper_chip_input(it._pix.data_ptr(), tab.cuda().data_ptr(), it.means.data_ptr(), sq.data_ptr(), B, S,
                                  torch.cuda.current_stream().cuda_stream))
    assert torch.equal(sq[:, :, :SH, :SW], data)
    info = b['im_info'].cpu().numpy()
    for k in range(B):
        dh, dw = int(info[k, 0]), int(info[k, 1])
        assert float(data[k, :, dh:, :].abs().sum()) == 0.0 and float(data[k, :, :, dw:].abs().sum()) == 0.0
        assert float(data[k, :, :dh, :dw].abs().max()) > 1.0


def test_autofocus_pyramid_runs_and_device_nms_equals_host_nms():
    import torch
    from sniper_b200 import model, synth_batch
    from sniper_b200 import tester as TS
    cfg = _small_cfg()
    rng = np.random.RandomState(2)
    db = [dict(width=W, height=H, image_data=rng.randint(0, 256, (H, W, 3)).astype(np.uint8))
          for (W, H) in ((320, 240), (240, 320), (300, 200))]
    mc = model.Cfg()
    mc.batch_images = 2
    net = model.SniperResNet101(mc, deform_offset_std=0.01)
    net.train_step(synth_batch.make_batch(2, seed=7, device="cuda"), lr=0.001)      # non-trivial moving statistics
    net.enable_autofocus(seed=3)
    w0 = net.P.w.clone()
    boxes_d, stats = TS.imdb_detection_wrapper(net, cfg, db, nms_backend="device")
    assert len(stats['scales']) == 3 and stats['scales'][0]['chips'] == 3
    for st in stats['scales'][:2]:
        assert 0.0 <= st['pixels_next_scale_pct'] <= 100.0 + 1e-9
    for r in db:                                     # chips of the last scale lie inside their image
        c = r['inference_crops']
        assert c.shape[1:] == (4,) or c.shape[0] == 0
        if len(c):
            assert (c[:, 0] >= 0).all() and (c[:, 1] >= 0).all() and (c[:, 2] <= r['width'] + 1e-6).all() \
                and (c[:, 3] <= r['height'] + 1e-6).all()
    chips_last = [np.array(r['inference_crops'], copy=True) for r in db]
    n = 0
    for j in range(1, 81):
        for i in range(3):
            d = boxes_d[j][i]
            assert d.ndim == 2 and d.shape[1] == 5 and np.isfinite(d).all()
            n += len(d)
    assert n <= 3 * cfg.TEST.MAX_PER_IMAGE + 80 * 3          # MAX_PER_IMAGE cut (ties may keep a few more)
    # same run, host soft-NMS: same detections (the network part is deterministic at inference)
    boxes_h, _ = TS.imdb_detection_wrapper(net, cfg, db, nms_backend="host")
    for r, c in zip(db, chips_last):
        assert np.array_equal(np.asarray(r['inference_crops']), c)
    for j in range(1, 81):
        for i in range(3):
            a, b = boxes_d[j][i], boxes_h[j][i]
            assert a.shape == b.shape and np.array_equal(a[:, :4], b[:, :4]) and np.abs(a[:, 4] - b[:, 4]).max(initial=0) < 1e-6
    assert torch.equal(w0, net.P.w)                  # inference touches no parameter
