"""FocusChip generation (lib/chips/chips_inference.py) restated without OpenCV: primitives against their definitions,
gmask on hand-checked maps, the cover / size / bounds properties the merge loop guarantees, add_chips bookkeeping and
the chip-border pruning of Tester.get_detections."""
import math
from types import SimpleNamespace

import numpy as np

from sniper_b200 import chips_inference as CI


def _dilate_ref(mask, d):
    H, W = mask.shape
    a = d // 2
    out = np.zeros_like(mask)
    for y in range(H):
        for x in range(W):
            m = None
            for j in range(d):
                for i in range(d):
                    yy, xx = y + j - a, x + i - a
                    if 0 <= yy < H and 0 <= xx < W:
                        m = mask[yy, xx] if m is None else max(m, mask[yy, xx])
            out[y, x] = m
    return out


def test_dilate_matches_definition():
    rng = np.random.RandomState(0)
    for d in (1, 2, 3, 4, 5):
        m = (rng.rand(13, 17) > 0.85).astype(np.float32)
        assert np.array_equal(CI.dilate(m, d), _dilate_ref(m, d))


def test_contour_rects_components_and_holes():
    m = np.zeros((12, 16), np.uint8)
    m[1:3, 1:4] = 255                     # blob A
    m[5:10, 6:12] = 255                   # ring B ...
    m[7, 8:10] = 0                        # ... with a 1 x 2 hole
    m[3, 4] = 255                         # touches A diagonally: same 8-connected component
    rects = CI.contour_rects(m)
    assert (1, 1, 4, 3) in rects          # A + the diagonal pixel: x 1..4, y 1..3
    assert (6, 5, 6, 5) in rects          # outer border of the ring
    assert (7, 6, 4, 3) in rects          # hole border: the hole's box (x 8..9, y 7) grown by one pixel
    assert len(rects) == 3
    # background connected to the frame is not a hole
    m2 = np.zeros((6, 6), np.uint8); m2[2:4, 2:4] = 255
    assert CI.contour_rects(m2) == [(2, 2, 2, 2)]
    assert CI.contour_rects(np.zeros((4, 4), np.uint8)) == []


def test_gmask_single_pixel_gives_one_min_size_chip():
    m = np.zeros((32, 40), np.float32)
    m[10, 20] = 0.9
    chips = CI.gmask(m, 3, 0.5, ms=8, im_width=40 * 16, im_height=32 * 16, cscale=2.0)
    # dilated blob = 3 x 3 cells centred at (20, 10): centre cell 20 -> chip cells [20 - 4, 20 + 4) x [10 - 4, 10 + 4)
    assert chips == [[16 * 16 / 2.0, 6 * 16 / 2.0, 24 * 16 / 2.0, 14 * 16 / 2.0]]


def test_gmask_corner_is_clamped_inside_the_map():
    m = np.zeros((20, 20), np.float32)
    m[0, 0] = 1.0
    m[19, 19] = 1.0
    chips = CI.gmask(m, 3, 0.5, ms=6, im_width=20 * 16 - 5, im_height=20 * 16, cscale=1.0)
    assert [0.0, 0.0, 96.0, 96.0] in chips
    # right edge: x2 clipped to the (non multiple of 16) width and x1 pulled back to keep ms * 16 px
    assert [20 * 16 - 5 - 96.0, 224.0, 20 * 16 - 5.0, 320.0] in chips
    assert len(chips) == 2


def test_gmask_properties_on_random_maps():
    rng = np.random.RandomState(3)
    for trial in range(20):
        H, W = rng.randint(12, 40), rng.randint(12, 50)
        m = rng.rand(H, W).astype(np.float32) ** 6          # sparse high values
        ms = int(rng.choice([4, 8, 16]))
        d = int(rng.choice([1, 3, 5]))
        thr = 0.3
        imw, imh = W * 16 - int(rng.randint(0, 16)), H * 16 - int(rng.randint(0, 16))
        chips = CI.gmask(m, d, thr, ms=ms, im_width=imw, im_height=imh, cscale=1.0)
        fg = CI.dilate((m >= thr).astype(np.float32), d) > 0
        iw, ih = int(math.ceil(imw / 16.0)), int(math.ceil(imh / 16.0))
        cover = np.zeros((H, W), bool)
        for x1, y1, x2, y2 in chips:
            assert 0 <= x1 < x2 <= imw and 0 <= y1 < y2 <= imh
            assert x2 - x1 >= min(ms * 16, imw) - 16 and y2 - y1 >= min(ms * 16, imh) - 16
            cover[int(y1) // 16:int(math.ceil(y2 / 16.0)), int(x1) // 16:int(math.ceil(x2 / 16.0))] = True
        assert not (fg[:ih, :iw] & ~cover[:ih, :iw]).any()   # every FocusPixel of the map lies in some chip
        # chips of the final iteration are the connected regions of the filled mask: pairwise disjoint in cell space
        cells = [(int(x1) // 16, int(y1) // 16, int(math.ceil(x2 / 16.0)), int(math.ceil(y2 / 16.0))) for x1, y1, x2, y2 in chips]
        assert len(chips) == len(set(map(tuple, chips)))
        assert len(cells) >= (1 if fg[:ih, :iw].any() else 0)


def test_add_chips_shifts_by_the_parent_crop_and_reports_areas():
    cfg = SimpleNamespace(TEST=SimpleNamespace(SCALES=[(480, 512), (800, 1280), (1400, 2000)],
                                               CHIP_HYPERPARAMS=[(3, 0.02, 16), (3, 0.2, 20)]))
    roidb = [dict(width=640, height=480, inference_crops=np.array([[0, 0, 640, 480]]))]
    cs = CI.image_scale(640, 480, cfg.TEST.SCALES[0])
    assert cs == 512.0 / 640.0                               # 480 / 480 = 1 would make the long side 640 > 512
    mh, mw = int(math.ceil(480 * cs / 16.0)), int(math.ceil(640 * cs / 16.0))
    fmap = np.zeros((mh, mw), np.float32)
    fmap[5, 7] = 0.5
    area = CI.add_chips(roidb, [[(None, fmap)]], 0, cfg)
    crops = roidb[0]['inference_crops']
    assert crops.shape == (1, 4)
    x1, y1, x2, y2 = crops[0]
    assert x1 >= 0 and y1 >= 0 and x2 <= 640 + 1e-9 and y2 <= 480 + 1e-9
    assert x1 <= 7 * 16 / cs <= x2 and y1 <= 5 * 16 / cs <= y2
    ts = CI.image_scale(640, 480, cfg.TEST.SCALES[1])
    assert abs(area[1] - 640 * 480 * ts * ts / 1e6) < 1e-9
    assert abs(area[0] - (x2 - x1) * (y2 - y1) * ts * ts / 1e6) < 1e-9
    # second level: chips are found inside the previous chip and shifted by its origin
    fmap2 = np.zeros((int(math.ceil((y2 - y1) * ts / 16.0)), int(math.ceil((x2 - x1) * ts / 16.0))), np.float32)
    fmap2[2, 3] = 1.0
    CI.add_chips(roidb, [[(None, fmap2)]], 1, cfg)
    c2 = roidb[0]['inference_crops'][0]
    assert c2[0] >= x1 - 1e-9 and c2[1] >= y1 - 1e-9


def test_project_and_prune_equals_the_per_detection_loop():
    rng = np.random.RandomState(5)
    W, H = 1000, 700
    for chip in ([0, 0, 1000, 700], [100, 50, 600, 400], [0, 200, 512, 700], [488, 0, 1000, 512]):
        dets = np.hstack([rng.rand(200, 2) * 200, 200 + rng.rand(200, 2) * 300, rng.rand(200, 1)])
        dets[:20, 0] = rng.rand(20) * 12                     # near the left chip border
        dets[20:40, 3] = (chip[3] - chip[1]) - rng.rand(20) * 12
        got = CI.project_and_prune(dets, chip, W, H)
        exp = []
        for d in dets:
            t = d.copy(); t[0] += chip[0]; t[2] += chip[0]; t[1] += chip[1]; t[3] += chip[1]
            if CI.check_valid(t, chip, W, H):
                exp.append(t)
        exp = np.array(exp) if exp else np.zeros((0, 5))
        assert got.shape == exp.shape and np.array_equal(got, exp)
    assert CI.project_and_prune(np.zeros((0, 5)), [0, 0, 10, 10], W, H).shape == (0, 5)
