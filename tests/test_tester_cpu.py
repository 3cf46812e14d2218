"""MNIteratorTestAutoFocus batching logic on the host (no canvas is built without a GPU): chips sorted by area,
horizontal group before the vertical one, groups padded by repeating their last chips, a batch padded to its largest
resized chip, crop / resize table rows as im_worker.worker_autofocus computes them."""
import math

import numpy as np

from sniper_b200 import chips_inference as CI
from sniper_b200 import iterator as IT
from sniper_b200 import tester as TS


def _roidb():
    rng = np.random.RandomState(0)
    db = []
    for (W, H) in ((640, 480), (480, 640), (500, 375)):
        db.append(dict(width=W, height=H, image_data=rng.randint(0, 256, (H, W, 3)).astype(np.uint8)))
    db[0]['inference_crops'] = np.array([[0, 0, 640, 480]])
    db[1]['inference_crops'] = np.array([[10.5, 20.25, 210.5, 420.0], [0, 0, 480, 640], [100, 100, 400, 300]])
    db[2]['inference_crops'] = np.array([[0, 0, 500, 375], [250.7, 0, 500, 200.2]])
    return db


def test_grouping_padding_and_table_rows():
    cfg = IT.default_config()
    db = _roidb()
    scale = (160, 256)
    it = TS.MNIteratorTestAutoFocus(db, cfg, scale, batch_size=4, device="cpu")
    assert it.n_chips == 6 and it.size == 8                    # 4 horizontal chips, 2 vertical ones padded to 4
    crops = [c for r in db for c in r['inference_crops']]
    w = np.array([c[2] - c[0] for c in crops]); h = np.array([c[3] - c[1] for c in crops])
    first, second = it.inds[:4], it.inds[4:]
    assert all(w[i] >= h[i] for i in first) and all(w[i] < h[i] for i in second)
    assert list((w * h)[first]) == sorted((w * h)[first])       # by area inside a group
    assert second[2] == second[1] and second[3] == second[1] or set(second[2:]) <= set(second[:2])   # padding repeats
    seen = []
    for batch in it:
        SH, SW = batch['canvas']
        assert SH % 32 == 0 and SW % 128 == 0
        for k in range(4):
            im_id, chip_id = batch['im_ids'][k], batch['chip_ids'][k]
            r = db[im_id]
            c = r['inference_crops'][chip_id]
            s = CI.image_scale(r['width'], r['height'], scale)
            assert abs(batch['scales'][k] - s) < 1e-15
            off, sh, sw, dh, dw, flip, bits, _ = batch['table'][k]
            y1, y2 = max(int(c[1]), 0), min(int(c[3]), r['height'])
            x1, x2 = max(int(c[0]), 0), min(int(c[2]), r['width'])
            assert (sh, sw) == (y2 - y1, x2 - x1) and flip == 0
            assert np.int64(bits).view(np.float64) == s
            assert (dh, dw) == (int(np.rint(sh * s)), int(np.rint(sw * s)))
            assert dh <= SH and dw <= SW and SH >= int(math.ceil((c[3] - c[1]) * s)) and SW >= int(math.ceil((c[2] - c[0]) * s))
            src = batch['pixels'][off:off + sh * sw * 3].reshape(sh, sw, 3)
            assert np.array_equal(src, r['image_data'][y1:y2, x1:x2])
            assert tuple(batch['im_info'][k].tolist()) == (float(dh), float(dw), np.float32(s))
            seen.append((im_id, chip_id))
    assert set(seen) == {(0, 0), (1, 0), (1, 1), (1, 2), (2, 0), (2, 1)}


def test_empty_and_scale_switch():
    cfg = IT.default_config()
    db = _roidb()
    for r in db:
        r['inference_crops'] = np.zeros((0, 4))
    it = TS.MNIteratorTestAutoFocus(db, cfg, (480, 512), batch_size=2, device="cpu")
    assert it.size == 0 and list(it) == []
    db[0]['inference_crops'] = np.array([[0, 0, 640, 480]])
    it.set_scale((800, 1280)); it.reset()
    b = next(it)
    assert it.size == 2 and b['canvas'] == (800, 1152) and b['im_ids'].tolist() == [0, 0]     # 480 * 1.6667 = 800, 640 * 1.6667 = 1067 -> 1152
