"""Host half of the inference post-processing: the vectorised `threshold_detections` / `aggregate` assembly equal the
per-class / per-image loops of the reference (lib/inference.py:291-301, 152-230), checked with the host soft-NMS."""
import numpy as np

from sniper_b200 import host
from sniper_b200 import inference as I


def test_threshold_detections_equals_the_per_class_loop():
    rng = np.random.RandomState(0)
    K = 81
    for R in (0, 1, 300):
        sc = rng.rand(R, K).astype(np.float32) ** 8
        bx = rng.rand(R, 4) * 500
        new = I.threshold_detections(sc, bx, K, 1e-3)
        assert len(new) == K and new[0].shape == (0, 5)
        for j in range(1, K):
            inds = np.where(sc[:, j] > 1e-3)[0]
            old = np.hstack((bx[inds, 0:4], sc[inds, j, np.newaxis])).astype(np.float32)
            assert new[j].dtype == np.float32 and np.array_equal(new[j], old)


def _loop_problems(scale_cls_dets, valid_ranges, num_images, num_classes):
    problems = []
    for i in range(num_images):
        for j in range(1, num_classes):
            agg = np.empty((0, 5), dtype=np.float32)
            for all_cls_dets, vr in zip(scale_cls_dets, valid_ranges):
                for cls_dets in all_cls_dets[j][i]:
                    cls_dets = I._valid_range_filter(np.asarray(cls_dets, np.float32).reshape(-1, 5), vr)
                    if cls_dets.shape[0] > 0:
                        agg = np.vstack((agg, cls_dets))
            problems.append(agg)
    return problems


def test_aggregate_equals_the_nested_loops():
    rng = np.random.RandomState(1)
    NI, NC = 3, 6
    scales = []
    for s in range(3):
        allb = [[[] for _ in range(NI)] for _ in range(NC)]
        for j in range(1, NC):
            for i in range(NI):
                for chip in range(rng.randint(0, 3)):
                    n = rng.randint(0, 6)
                    b = rng.rand(n, 5).astype(np.float32)
                    b[:, :2] *= 100
                    b[:, 2:4] = b[:, :2] + rng.rand(n, 2) * 150
                    allb[j][i].append(b)
        scales.append(allb)
    vr = [(40, -1), (16, 90), (-1, 40)]
    ref = _loop_problems(scales, vr, NI, NC)
    out = I.aggregate(scales, vr, NI, NC, max_per_image=-1, backend="host")
    for i in range(NI):
        for j in range(1, NC):
            p = ref[i * (NC - 1) + (j - 1)]
            exp = host.cpu_soft_nms(p.copy(), sigma=0.55, Nt=0.3, threshold=0.001, method=2) if len(p) else p
            assert np.array_equal(out[j][i], exp), (i, j)
    # MAX_PER_IMAGE cut (:213-221): the image keeps its highest-scoring rows over all classes
    cut = I.aggregate(scales, vr, NI, NC, max_per_image=5, backend="host")
    for i in range(NI):
        allsc = np.hstack([out[j][i][:, 4] for j in range(1, NC)])
        kept = np.hstack([cut[j][i][:, 4] for j in range(1, NC)])
        if len(allsc) > 5:
            th = np.sort(allsc)[-5]
            assert np.array_equal(np.sort(kept), np.sort(allsc[allsc >= th]))
        else:
            assert np.array_equal(np.sort(kept), np.sort(allsc))
    empty = [[[[] for _ in range(NI)] for _ in range(NC)]]
    e = I.aggregate(empty, [(-1, -1)], NI, NC, backend="host")
    assert all(e[j][i].shape == (0, 5) for j in range(1, NC) for i in range(NI))


def test_threshold_with_pruning_equals_per_class_project_and_prune():
    from sniper_b200 import chips_inference as CI
    rng = np.random.RandomState(3)
    K = 81
    for chip in ([0, 0, 1000, 700], [100, 50, 600, 400], [488, 0, 1000, 512]):
        sc = rng.rand(300, K).astype(np.float32) ** 6
        bx = np.hstack([rng.rand(300, 2) * 200, 200 + rng.rand(300, 2) * 300])
        bx[:30, 0] = rng.rand(30) * 12                       # near the left chip border
        new = I.threshold_detections(sc, bx, K, 1e-3, prune=(chip, 1000, 700))
        base = I.threshold_detections(sc, bx, K, 1e-3)
        for j in range(1, K):
            exp = CI.project_and_prune(base[j], chip, 1000, 700)
            assert new[j].shape == exp.shape and np.array_equal(new[j], exp)
