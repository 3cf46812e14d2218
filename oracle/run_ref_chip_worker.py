"""TEST INFRASTRUCTURE ONLY.  Runs the reference's OWN `chip_worker` (lib/data_utils/data_workers.py:373-594:
chip_extractor, box_assigner) with its own `chip_generator` (lib/chips/chip_generator.py), `clip_boxes` /
`ignore_overlaps` (lib/bbox/bbox_transform.py over the reference's Cython bbox.pyx, oracle/_ref/ref_bbox) and its own
greedy cover (lib/chips/cchips.cpp compiled as it lies into oracle/_ref/libref_chips.so) in THIS container, to pin
sniper_b200/chip_worker.py and to produce tests/golden/chip_worker_ref.npz.

The sources are read as text from /root/reference; the class is cut out of data_workers.py and `np.float` (removed from
numpy) is rewritten to `float` in memory.  Nothing is copied into the repository.  The reference never seeds the C
rand() stream that cchips.cpp's random_shuffle draws from; callers seed it (libc srand) before each run so that the
reference and the product consume identical streams."""
import ctypes
import math
import os
import re
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
_libc = ctypes.CDLL(None)


def srand(seed):
    _libc.srand(ctypes.c_uint(seed))


def _ref_chips_module():
    L = ctypes.CDLL(os.path.join(HERE, "_ref", "libref_chips.so"))
    L.ref_chips_generate_noseed.restype = ctypes.c_int

    def generate(boxes, width, height, chipsize, stride):
        assert boxes.dtype == np.float32
        out = np.zeros((4096, 4), np.float32)
        n = L.ref_chips_generate_noseed(boxes.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(boxes.shape[0]),
                                        ctypes.c_int(width), ctypes.c_int(height), ctypes.c_int(chipsize),
                                        ctypes.c_int(stride), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(4096))
        return out[:n].tolist()            # chips.pyx returns vector<vector<float>> -> list of lists of Python floats
    return types.SimpleNamespace(generate=generate)


def load_reference_chip_worker(ref=REF):
    sys.path.insert(0, os.path.join(HERE, "_ref"))
    import ref_bbox
    bt = open(os.path.join(ref, "lib", "bbox", "bbox_transform.py")).read()
    bt = bt.replace("from bbox import bbox_overlaps_cython, ignore_overlaps_cython", "")
    bt = re.sub(r"np\.float\b", "float", bt)
    ns_bt = {"bbox_overlaps_cython": ref_bbox.bbox_overlaps_cython, "ignore_overlaps_cython": ref_bbox.ignore_overlaps_cython}
    exec(compile(bt, "bbox_transform.py", "exec"), ns_bt)
    cg = open(os.path.join(ref, "lib", "chips", "chip_generator.py")).read()
    cg = cg.replace("import chips\n", "").replace("from bbox.bbox_transform import clip_boxes, ignore_overlaps\n", "")
    ns_cg = {"chips": _ref_chips_module(), "clip_boxes": ns_bt["clip_boxes"], "ignore_overlaps": ns_bt["ignore_overlaps"]}
    exec(compile(cg, "chip_generator.py", "exec"), ns_cg)
    dw = open(os.path.join(ref, "lib", "data_utils", "data_workers.py")).read()
    start = dw.index("class chip_worker(object):")
    nxt = dw.find("\nclass ", start + 10)
    cls = re.sub(r"np\.float\b", "float", dw[start:nxt if nxt > 0 else len(dw)])
    ns = {"np": np, "math": math, "chip_generator": ns_cg["chip_generator"], "ignore_overlaps": ns_bt["ignore_overlaps"]}
    exec(compile(cls, "data_workers.py[chip_worker]", "exec"), ns)
    return ns["chip_worker"]


def make_cfg(scales=((1400, 2000), (800, 1280), (-1, 512)), valid_ranges=((-1, 80), (32, 150), (120, -1)), neg=True):
    """TRAIN section of configs/faster/sniper_res101_e2e.yml:76-78, 99-101 (resolution mode), or float factors."""
    S = types.SimpleNamespace
    sc = [tuple(s) if isinstance(s, (tuple, list)) else float(s) for s in scales]
    return S(TRAIN=S(VALID_RANGES=[tuple(v) for v in valid_ranges], SCALES=sc, CPP_CHIPS=True, USE_NEG_CHIPS=neg))


def synth_roidb(seed, width=1333, height=800, n_gt=20, n_prop=2000):
    """sqrt(area) log-uniform in [8,400], aspect in [0.5,2], centres uniform, clipped to the image;
    GT rows carry max_overlaps == 1, proposals < 1 (the roidb convention chip_extractor relies on)."""
    rng = np.random.RandomState(seed)
    n = n_gt + n_prop
    s = np.exp(rng.uniform(np.log(8), np.log(400), n))
    ar = np.exp(rng.uniform(np.log(0.5), np.log(2.0), n))
    w, h = s * np.sqrt(ar), s / np.sqrt(ar)
    cx, cy = rng.uniform(0, width, n), rng.uniform(0, height, n)
    boxes = np.stack([np.clip(cx - w / 2, 0, width - 1), np.clip(cy - h / 2, 0, height - 1),
                      np.clip(cx + w / 2, 0, width - 1), np.clip(cy + h / 2, 0, height - 1)], 1).astype(np.float32)
    mo = np.concatenate([np.ones(n_gt), rng.uniform(0, 0.9, n_prop)]).astype(np.float32)
    return {"width": width, "height": height, "boxes": boxes, "max_overlaps": mo}


def run(worker_cls, cfg, roidb, seed, stride=58, chip_size=512):
    """chip_extractor + box_assigner of `worker_cls` on a copy of `roidb` with a fixed chip stride and rand() seed."""
    np.random.seed(0)
    w = worker_cls(cfg, chip_size)
    w.chip_stride = stride
    w.chip_generator.chip_stride = stride
    r = dict(roidb)
    srand(seed)
    r["crops"] = w.chip_extractor(r)
    out = w.box_assigner(r)
    return r["crops"], out


if __name__ == "__main__":
    W = load_reference_chip_worker()
    crops, out = run(W, make_cfg(), synth_roidb(0), seed=1)
    print(len(crops), "positive chips;", len(out[1]), "negative chips;", [len(p) for p in out[0]][:8])
