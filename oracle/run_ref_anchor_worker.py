"""TEST INFRASTRUCTURE ONLY.  Runs the reference's OWN RPN anchor matching -- `anchor_worker.worker`
(lib/data_utils/data_workers.py:130-371) with its helpers `generate_anchors` (lib/data_utils/generate_anchor.py) and
`bbox_transform` / `clip_boxes` / `filter_boxes` (lib/bbox/bbox_transform.py) and the reference's Cython
`bbox_overlaps_cython` (oracle/_ref/ref_bbox, see build_ref_cython.py) -- in THIS container, to validate
oracle/anchor_target_np.py and to produce the golden vectors of tests/golden/anchor_target_ref.npz.

The reference is Python 2 and imports mxnet / cv2 at module level, so the sources are read as text from
/root/reference, the class `anchor_worker` is cut out of data_workers.py, and the Python-2-only tokens are rewritten
in memory (nothing is copied into the repo):
    xrange -> range;   np.float -> float;   `chip_size / cfg.network.RPN_FEAT_STRIDE` -> `//` (Py2 integer division);
    `mx.nd.array(x, ...)` -> numpy array (the worker only wraps its results in NDArrays).
The arithmetic and the control flow are the reference's."""
import os
import re
import sys
import types

import numpy as np

REF = "/root/reference"


def load_reference_worker(ref=REF):
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "_ref"))
    import ref_bbox                                             # the reference's bbox.pyx, cythonized
    ga = open(os.path.join(ref, "lib", "data_utils", "generate_anchor.py")).read().replace("xrange", "range")
    ns_ga = {}
    exec(compile(ga, "generate_anchor.py", "exec"), ns_ga)
    bt = open(os.path.join(ref, "lib", "bbox", "bbox_transform.py")).read()
    bt = bt.replace("from bbox import bbox_overlaps_cython, ignore_overlaps_cython", "")
    bt = re.sub(r"np\.float\b", "float", bt)
    ns_bt = {"bbox_overlaps_cython": ref_bbox.bbox_overlaps_cython, "ignore_overlaps_cython": ref_bbox.ignore_overlaps_cython}
    exec(compile(bt, "bbox_transform.py", "exec"), ns_bt)
    dw = open(os.path.join(ref, "lib", "data_utils", "data_workers.py")).read()
    start = dw.index("class anchor_worker(object):")
    nxt = dw.find("\nclass ", start + 10)
    cls = dw[start:nxt if nxt > 0 else len(dw)]
    cls = cls.replace("chip_size / cfg.network.RPN_FEAT_STRIDE", "chip_size // cfg.network.RPN_FEAT_STRIDE")
    cls = re.sub(r"np\.float\b", "float", cls)
    mx = types.SimpleNamespace(nd=types.SimpleNamespace(array=lambda x, dtype=None: np.asarray(x)))
    ns = dict(ns_bt)
    ns.update({"np": np, "npr": np.random, "math": __import__("math"), "mx": mx,
               "generate_anchors": ns_ga["generate_anchors"]})
    exec(compile(cls, "data_workers.py[anchor_worker]", "exec"), ns)
    return ns["anchor_worker"]


def make_cfg(scales=(2, 4, 7, 10, 13, 16, 24), ratios=(0.5, 1, 2), stride=16):
    S = types.SimpleNamespace
    return S(network=S(ANCHOR_SCALES=list(scales), ANCHOR_RATIOS=list(ratios), RPN_FEAT_STRIDE=stride),
             TRAIN=S(AUTO_FOCUS=False, AUTO_FOCUS_DC_LOW=0, AUTO_FOCUS_DC_HIGH=0, AUTO_FOCUS_SMALL_THRESH=0,
                     RPN_BATCH_SIZE=256, RPN_POSITIVE_OVERLAP=0.5, RPN_NEGATIVE_OVERLAP=0.4, RPN_FG_FRACTION=0.5))


def synth_case(seed, n_gt, n_valid, chip=512):
    """Integer-valued boxes inside the chip (>= 10 px), the first n_valid of them 'valid' for this chip."""
    rng = np.random.RandomState(seed)
    w = rng.randint(12, 260, n_gt).astype(np.float64)
    h = rng.randint(12, 260, n_gt).astype(np.float64)
    x1 = np.floor(rng.uniform(0, chip - 1 - w)).clip(0)
    y1 = np.floor(rng.uniform(0, chip - 1 - h)).clip(0)
    boxes = np.stack([x1, y1, x1 + w, y1 + h], 1)
    classes = rng.randint(1, 81, (n_gt, 1)).astype(np.float64)
    return boxes, classes, np.arange(n_gt), np.arange(n_valid)


def run_reference(worker, boxes, classes, gtids, nids, seed, chip=512):
    np.random.seed(seed)                                        # the worker draws its subsamples from numpy.random
    im_info = np.array([chip, chip, 1.0])
    out = worker.worker([im_info, np.array([0.0, 0.0]), 1.0, nids.copy(), gtids.copy(), boxes[gtids].copy(), boxes.copy(),
                         classes.copy()])
    labels, targets_pos, pids, fgt = out[:4]
    return (np.asarray(labels, np.float32).ravel(), np.asarray(targets_pos, np.float32), np.asarray(pids).astype(np.int64),
            np.asarray(fgt, np.float64))


if __name__ == "__main__":
    W = load_reference_worker()(make_cfg(), 512)
    b, c, g, n = synth_case(0, 12, 9)
    lab, tg, pids, fgt = run_reference(W, b, c, g, n, seed=3)
    print("labels fg/bg/ignore:", int((lab == 1).sum()), int((lab == 0).sum()), int((lab == -1).sum()), "targets", tg.shape)
