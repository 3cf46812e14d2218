"""AutoFocus inference chips: FocusPixel map -> FocusChips of the next scale, and pruning of chip-border detections.

Host-side mirror of `lib/chips/chips_inference.py` (`gmask` :12-89, `add_chips` :91-173) and of the projection /
`check_valid` step of `Tester.get_detections` (lib/inference.py:235-256, 335-351).  Same function names, argument meaning
and return values as the reference.

The reference leans on OpenCV for three primitives (this image has no OpenCV, so parity of those three is UNPINNED and
restated here from the OpenCV documentation):
  * `cv2.dilate(mask, ones((d, d)))`: max over the d x d window anchored at (d // 2, d // 2), pixels outside ignored;
  * `cv2.findContours(mask, RETR_LIST, ...)`: every outer border (one per 8-connected foreground component) AND every hole
    border (one per background region enclosed by foreground; a hole border runs over the foreground pixels around it);
  * `cv2.boundingRect(contour)`: tight box of the border pixels = the component's box for an outer border, the hole's box
    grown by one pixel for a hole border.
Contours are listed last-found-first (raster scan of the start pixel, the order `RETR_LIST` yields).  Everything else --
the integer arithmetic of the merge loop (Python 2 `/` on ints = floor division), the iteration until the chip count
stops changing, the rescaling -- follows the reference line by line.
"""
import math

import numpy as np
from scipy import ndimage

_S8 = np.ones((3, 3), dtype=bool)
_S4 = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], dtype=bool)


def dilate(mask, d):
    """cv2.dilate(mask, np.ones((d, d), np.uint8)) for a 2-D array: dst(y, x) = max src(y + j - a, x + i - a),
    0 <= i, j < d, a = d // 2, positions outside the image skipped."""
    H, W = mask.shape
    a = d // 2
    out = mask.copy()
    for j in range(d):
        for i in range(d):
            dy, dx = j - a, i - a
            ys0, ys1 = max(0, dy), min(H, H + dy)
            xs0, xs1 = max(0, dx), min(W, W + dx)
            if ys0 >= ys1 or xs0 >= xs1:
                continue
            out[ys0 - dy:ys1 - dy, xs0 - dx:xs1 - dx] = np.maximum(out[ys0 - dy:ys1 - dy, xs0 - dx:xs1 - dx],
                                                                   mask[ys0:ys1, xs0:xs1])
    return out


def contour_rects(mask):
    """Bounding rectangles (x, y, w, h) of the borders `cv2.findContours(mask, cv2.RETR_LIST, ...)` returns."""
    fg = mask != 0
    H, W = fg.shape
    found = []                                   # (start_y, start_x, rect)
    lab, n = ndimage.label(fg, structure=_S8)
    for k, sl in enumerate(ndimage.find_objects(lab)):
        ys, xs = sl
        row = np.where(lab[ys.start, xs] == k + 1)[0]
        found.append((ys.start, xs.start + int(row[0]), (xs.start, ys.start, xs.stop - xs.start, ys.stop - ys.start)))
    blab, bn = ndimage.label(~fg, structure=_S4)
    if bn:
        border = set(np.unique(np.concatenate([blab[0, :], blab[-1, :], blab[:, 0], blab[:, -1]])).tolist())
        for k, sl in enumerate(ndimage.find_objects(blab)):
            if (k + 1) in border:
                continue                          # connected to the frame: not a hole
            ys, xs = sl
            x0, y0 = max(xs.start - 1, 0), max(ys.start - 1, 0)
            x1, y1 = min(xs.stop + 1, W), min(ys.stop + 1, H)
            # the border of a hole starts at the foreground pixel left of the hole's first pixel in scan order
            row = np.where(blab[ys.start, xs] == k + 1)[0]
            found.append((ys.start, xs.start + int(row[0]) - 1, (x0, y0, x1 - x0, y1 - y0)))
    found.sort(key=lambda t: (t[0], t[1]), reverse=True)
    return [r for _, _, r in found]


def _place(x, y, w, h, ms, iw, ih):
    """One contour rectangle -> chip rectangle of at least ms x ms cells kept inside the iw x ih map (gmask :30-48)."""
    cx = (x + x + w) // 2
    cy = (y + y + h) // 2
    w = max(ms, w)
    h = max(ms, h)
    if cx + w // 2 >= iw:
        x = iw - w if iw - w >= 0 else 0
    elif cx - w // 2 < 0:
        x = 0
    else:
        x = cx - w // 2
    if cy + h // 2 >= ih:
        y = ih - h if ih - h >= 0 else 0
    elif cy - h // 2 < 0:
        y = 0
    else:
        y = cy - h // 2
    return x, y, w, h


def gmask(mask, d, thresh_value=0.5, ms=16, im_width=0, im_height=0, cscale=1):
    """FocusPixel probability map (stride 16) -> list of FocusChips [x1, y1, x2, y2] in the coordinates of the
    original image (chips_inference.py:12-89)."""
    mask = np.array(mask, copy=True)
    iw = int(math.ceil(float(im_width) / 16))
    ih = int(math.ceil(float(im_height) / 16))
    hi = mask >= thresh_value
    mask[hi] = 1
    mask[~hi] = 0
    mask = dilate(mask, d)
    mask *= 255
    rects = contour_rects(mask.astype(np.uint8))
    chips = []
    nchips = -1
    while nchips != len(chips):
        nchips = len(chips)
        chips = []
        for (x, y, w, h) in rects:
            x, y, w, h = _place(x, y, w, h, ms, iw, ih)
            mask[y:y + h, x:x + w] = 255
        rects = contour_rects(mask.astype(np.uint8))
        for (x, y, w, h) in rects:
            x, y, w, h = _place(x, y, w, h, ms, iw, ih)
            chips.append([x, y, x + w, y + h])
    schips = []
    for c in chips:
        x1, y1, x2, y2 = c[0] * 16, c[1] * 16, c[2] * 16, c[3] * 16
        if x2 > im_width:
            x2 = im_width
            x1 = max(min(x1, x2 - ms * 16), 0)
        if y2 > im_height:
            y2 = im_height
            y1 = max(min(y1, y2 - ms * 16), 0)
        schips.append([x1 / cscale, y1 / cscale, x2 / cscale, y2 / cscale])
    return schips


def image_scale(width, height, target):
    """The resize factor the test iterator applies for TEST.SCALES entry `target` = (min side, max side)."""
    im_size_min, im_size_max = min(width, height), max(width, height)
    s = float(target[0]) / float(im_size_min)
    if np.round(s * im_size_max) > target[1]:
        s = float(target[1]) / float(im_size_max)
    return s


def add_chips(roidb, maps, scale_id, cfg):
    """Replaces every image's `inference_crops` by the FocusChips found at scale `scale_id` (to be processed at scale
    scale_id + 1).  maps[i][j] = (anything, FocusPixel map of crop j of image i).  Returns [chip_area, total_area] in
    megapixels at the next scale (chips_inference.py:91-173)."""
    total_area = 0
    chip_area = 0
    hp = cfg.TEST.CHIP_HYPERPARAMS[scale_id]
    for i, r in enumerate(roidb):
        cur_chips = []
        im_width, im_height = r['width'], r['height']
        cscale = image_scale(im_width, im_height, cfg.TEST.SCALES[scale_id])
        tcscale = image_scale(im_width, im_height, cfg.TEST.SCALES[scale_id + 1])
        total_area = total_area + (im_width * im_height * tcscale * tcscale) / (1000. * 1000.)
        for j in range(len(maps[i])):
            cmap = maps[i][j][1]
            cur_crop = r['inference_crops'][j]
            crop_width = cur_crop[2] - cur_crop[0]
            crop_height = cur_crop[3] - cur_crop[1]
            chips = gmask(cmap, hp[0], hp[1], ms=hp[2], im_width=crop_width * cscale, im_height=crop_height * cscale,
                          cscale=cscale)
            for c in chips:
                c[0] += cur_crop[0]
                c[1] += cur_crop[1]
                c[2] += cur_crop[0]
                c[3] += cur_crop[1]
            for c in chips:
                tarea = (c[2] - c[0]) * (c[3] - c[1]) * tcscale * tcscale
                chip_area = chip_area + tarea / (1000. * 1000.)
            cur_chips += chips
        roidb[i]['inference_crops'] = np.array(cur_chips)
    return [chip_area, total_area]


def check_valid(det, chip, im_width, im_height, delta=10):
    """A detection within `delta` px of a chip border that is not an image border is cut off by the chip: drop it
    (Tester.get_detections, lib/inference.py:235-256)."""
    dx1, dy1, dx2, dy2 = det[0], det[1], det[2], det[3]
    cx1, cy1, cx2, cy2 = chip[0], chip[1], chip[2], chip[3]
    if cx1 >= 0.5 and abs(dx1 - cx1) < delta:
        return False
    if cy1 >= 0.5 and abs(dy1 - cy1) < delta:
        return False
    if cx2 < im_width - 0.5 and abs(dx2 - cx2) < delta:
        return False
    if cy2 < im_height - 0.5 and abs(dy2 - cy2) < delta:
        return False
    return True


def prune_mask(d, chip, im_width, im_height, delta=10):
    """check_valid for every row of d ([n, >=4], image coordinates) at once."""
    keep = np.ones(d.shape[0], dtype=bool)
    if chip[0] >= 0.5:
        keep &= ~(np.abs(d[:, 0] - chip[0]) < delta)
    if chip[1] >= 0.5:
        keep &= ~(np.abs(d[:, 1] - chip[1]) < delta)
    if chip[2] < im_width - 0.5:
        keep &= ~(np.abs(d[:, 2] - chip[2]) < delta)
    if chip[3] < im_height - 0.5:
        keep &= ~(np.abs(d[:, 3] - chip[3]) < delta)
    return keep


def project_and_prune(cls_dets, chip, im_width, im_height, delta=10):
    """Detections of one chip (chip coordinates, rows x1 y1 x2 y2 score) -> image coordinates, border-cut ones removed
    (`do_pruning` branch, lib/inference.py:335-351).  Vectorised form of the per-detection loop."""
    d = np.array(cls_dets, dtype=np.float64, copy=True).reshape(-1, 5)
    if d.shape[0] == 0:
        return np.zeros((0, 5))
    d[:, 0] += chip[0]; d[:, 2] += chip[0]
    d[:, 1] += chip[1]; d[:, 3] += chip[1]
    d = d[prune_mask(d, chip, im_width, im_height, delta)]
    return d if d.shape[0] > 0 else np.zeros((0, 5))
