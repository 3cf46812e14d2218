"""Generates the committed golden fixtures (run in the build container, where /root/reference exists).

  chips_config1.npz : output of the REFERENCE's own lib/chips/cchips.cpp (compiled as it lies into
                      oracle/_ref/libref_chips.so by oracle/Makefile) on BASELINE config 1: one 1333x800 image,
                      20 GT boxes + 400 proposals, three scales, chip 512, fixed stride, srand(seed).
  mpt_small.npz     : oracle/mpt.c on a seeded 2-chip input (the reference ships no vectors for this operator).
  psroi_small.npz   : oracle/psroi.c on the reference's own test shapes (test_operator.py:4358-4389).
  host_refcython.npz / anchor_target_ref.npz : outputs of the reference's own Cython host code (cpu_nms, cpu_soft_nms,
                      bbox overlaps; oracle/build_ref_cython.py) and of its anchor_worker.worker (oracle/run_ref_anchor_worker.py).
  mpt_refcpu.npz    : output of the REFERENCE's own CPU operators -- MultiProposalTargetOp<cpu>::Forward
                      (multi_proposal_target.cc) and MultiProposalGPUOp<cpu>::Forward (multi_proposal.cc), compiled as
                      they lie into oracle/_ref/libref_mpt.so / libref_mp.so -- on a seeded 1-chip input with tie-free
                      scores (inputs are regenerated from the seed by the test).
Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from sniper_b200 import synth  # noqa: E402


def config1_boxes(seed, n, W=1333, H=800):
    rng = np.random.RandomState(seed)
    s = np.exp(rng.uniform(np.log(8), np.log(400), n))
    ar = np.exp(rng.uniform(np.log(0.5), np.log(2.0), n))
    w, h = s * np.sqrt(ar), s / np.sqrt(ar)
    cx, cy = rng.uniform(0, W, n), rng.uniform(0, H, n)
    return np.stack([np.clip(cx - w / 2, 0, W - 1), np.clip(cy - h / 2, 0, H - 1), np.clip(cx + w / 2, 0, W - 1),
                     np.clip(cy + h / 2, 0, H - 1)], 1).astype(np.float32)


def main():
    assert O.ref_chips() is not None, "build oracle/_ref first (make -C oracle ref)"
    out = {}
    boxes0 = np.concatenate([config1_boxes(0, 20), config1_boxes(1, 400)])
    for k, (scale, stride, seed) in enumerate([(3.0, 58, 1), (1.667, 58, 1), (0.384, 58, 1)]):
        W, H = int(1333 * scale), int(800 * scale)
        b = boxes0 * np.float32(scale)
        b[:, [0, 2]] = np.clip(b[:, [0, 2]], 0, W - 1)
        b[:, [1, 3]] = np.clip(b[:, [1, 3]], 0, H - 1)
        out["boxes%d" % k] = b
        out["meta%d" % k] = np.array([W, H, 512, stride, seed], np.int64)
        out["chips%d" % k] = O.ref_chips_generate(b, W, H, 512, stride, seed=seed)
    np.savez_compressed(os.path.join(HERE, "chips_config1.npz"), **out)

    inp = synth.mpt_inputs(1234, 2, 21, 16, 16)
    res = O.multi_proposal_target(*inp)
    np.savez_compressed(os.path.join(HERE, "mpt_small.npz"), cls_prob=inp[0], bbox_pred=inp[1], im_info=inp[2],
                        gt_boxes=inp[3], valid_ranges=inp[4], rois=res["rois"], label=res["label"],
                        bbox_target=res["bbox_target"], bbox_weight=res["bbox_weight"], keep_idx=res["keep_idx"],
                        num_kept=res["num_kept"])

    rng = np.random.RandomState(3)
    data = rng.rand(1, 18, 14, 14).astype(np.float32)
    rois = np.array([[0, 10, 22, 161, 173], [0, 20, 15, 154, 160]], np.float32)
    trans = (rng.rand(2, 4, 3, 3).astype(np.float32) - 0.5)
    o, c, si = O.deform_psroi_fwd(data, rois, trans, 0.0625, 2, 3, 3, 3, 4, 0.1, False)
    o2, bins = O.psroi_fwd(data, rois, 0.0625, 2, 3, 3)
    np.savez_compressed(os.path.join(HERE, "psroi_small.npz"), data=data, rois=rois, trans=trans, deform_out=o,
                        deform_count=c, deform_sample_idx=si, psroi_out=o2, psroi_bins=bins)
    # reference CPU operators (the binaries built from /root/reference)
    seed, B = 77, 1
    cls_prob, bbox_pred, im_info, gts, vr = synth.mpt_inputs(seed, B)
    n = cls_prob[:, 21:].size
    fg = ((np.random.RandomState(1000 + seed).permutation(n) + 1.0) / (n + 1.0)).astype(np.float32).reshape(B, 21, 32, 32)
    cls_prob = cls_prob.copy()
    cls_prob[:, 21:] = fg
    cls_prob[:, :21] = 1.0 - fg
    ref = O.ref_multi_proposal_target(cls_prob, bbox_pred, im_info, gts, vr, bbox_scale=1.0)
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
    prois, pscores = O.ref_multi_proposal(cls_prob, bbox_pred, im_info)
    kept = int((pscores > 0).sum())
    np.savez_compressed(os.path.join(HERE, "mpt_refcpu.npz"), seed=np.int64(seed), fg=fg, rois=ref["rois"],
                        label=ref["label"], bbox_target=ref["bbox_target"], bbox_weight=ref["bbox_weight"],
                        proposal_rois=prois[:kept], proposal_scores=pscores[:kept])
    # reference Cython host code (oracle/_ref/ref_cpu_nms, ref_bbox built by oracle/build_ref_cython.py)
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import ref_bbox
    import ref_cpu_nms
    rng = np.random.RandomState(5)
    n = 500
    r = synth.rois_for_pool(rng, n, 1)
    dets = np.concatenate([r[:, 1:], (rng.permutation(n).reshape(-1, 1) + 1.0) / (n + 1.0)], 1).astype(np.float32)
    keep = np.array(ref_cpu_nms.cpu_nms(dets.copy(), 0.7), np.int32)
    soft = {m: np.array(ref_cpu_nms.cpu_soft_nms(dets.copy(), sigma=0.55, Nt=0.3, threshold=0.001, method=m)) for m in (0, 1, 2)}
    qa = synth.rois_for_pool(rng, 60, 1)[:, 1:].astype(np.float64)
    ov = ref_bbox.bbox_overlaps_cython(dets[:, :4].astype(np.float64), qa)
    np.savez_compressed(os.path.join(HERE, "host_refcython.npz"), dets=dets, keep=keep, soft0=soft[0], soft1=soft[1],
                        soft2=soft[2], query=qa, overlaps=ov)
    # the reference's own anchor_worker.worker (oracle/run_ref_anchor_worker.py)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import run_ref_anchor_worker as R
    worker = R.load_reference_worker()(R.make_cfg(), 512)
    boxes, classes, gtids, nids = R.synth_case(4, 60, 41)
    lab, tg, pids, fgt = R.run_reference(worker, boxes, classes, gtids, nids, seed=104)
    np.savez_compressed(os.path.join(HERE, "anchor_target_ref.npz"), boxes=boxes, n_valid=np.int64(41), seed=np.int64(104),
                        labels=lab, targets_pos=tg, pids=pids)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
