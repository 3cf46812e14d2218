"""Batched device soft-NMS against the host `cpu_soft_nms` of this repo, which is itself bit-identical to the reference's
Cython (tests/test_host_cpu.py, golden host_refcython.npz): same surviving boxes in the same order, scores to 1 float ulp
(the Gaussian weight goes through a double-precision exp on both sides: glibc on the host, CUDA libm on the device)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _problem(rng, n, W=1333, H=800):
    s = np.exp(rng.uniform(np.log(8), np.log(400), n))
    ar = np.exp(rng.uniform(np.log(0.5), np.log(2.0), n))
    w, h = s * np.sqrt(ar), s / np.sqrt(ar)
    # clustered centres so that boxes really overlap
    k = max(1, n // 12)
    cx = rng.uniform(0, W, k)[rng.randint(0, k, n)] + rng.randn(n) * 15
    cy = rng.uniform(0, H, k)[rng.randint(0, k, n)] + rng.randn(n) * 15
    d = np.stack([np.clip(cx - w / 2, 0, W - 1), np.clip(cy - h / 2, 0, H - 1), np.clip(cx + w / 2, 0, W - 1),
                  np.clip(cy + h / 2, 0, H - 1), rng.uniform(0.001, 1.0, n)], 1).astype(np.float32)
    return d


@pytest.mark.parametrize("method,sigma,Nt", [(2, 0.55, 0.3), (1, 0.5, 0.3), (3, 0.5, 0.45)])
def test_soft_nms_batched_matches_host(method, sigma, Nt):
    import torch
    from sniper_b200 import host, ops
    rng = np.random.RandomState(method)
    sizes = [0, 1, 2, 37, 300, 1000, 5, 640, 0, 129]
    probs = [_problem(rng, n) for n in sizes]
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    dets = torch.from_numpy(np.concatenate(probs, 0)).cuda()
    out, counts = ops.soft_nms_batched(dets, torch.from_numpy(offsets).cuda(), sigma=sigma, Nt=Nt, threshold=0.001, method=method)
    out, counts = out.cpu().numpy(), counts.cpu().numpy()
    for p, d in enumerate(probs):
        ref = host.cpu_soft_nms(d.copy(), sigma=sigma, Nt=Nt, threshold=0.001, method=method) if len(d) else d
        assert counts[p] == len(ref), (p, counts[p], len(ref))
        got = out[offsets[p]:offsets[p] + counts[p]]
        assert np.array_equal(got[:, :4], ref[:, :4]), p
        ulp = np.abs(got[:, 4].view(np.int32).astype(np.int64) - ref[:, 4].view(np.int32).astype(np.int64))
        assert ulp.max(initial=0) <= 1, (p, ulp.max())
    assert counts[5] < sizes[5] or method == 1      # among 1000 clustered boxes something is suppressed / decayed away


def test_aggregate_device_equals_host_path():
    """Tester.aggregate (lib/inference.py:152-230) with all (image, class) problems in one device launch == the same
    function through the host cpu_soft_nms (pinned by the reference's Cython)."""
    from sniper_b200 import inference
    rng = np.random.RandomState(5)
    num_images, num_classes = 3, 6
    valid_ranges = [(-1, 80), (32, 150), (120, -1)]
    scale_cls_dets = []
    for s in range(3):
        per_class = [[[] for _ in range(num_images)] for _ in range(num_classes)]
        for j in range(1, num_classes):
            for i in range(num_images):
                per_class[j][i] = [_problem(rng, int(rng.randint(0, 120)), 640, 480) for _ in range(int(rng.randint(1, 4)))]
        scale_cls_dets.append(per_class)
    dev = inference.aggregate(scale_cls_dets, valid_ranges, num_images, num_classes, max_per_image=40)
    hst = inference.aggregate(scale_cls_dets, valid_ranges, num_images, num_classes, max_per_image=40, backend="host")
    n = 0
    for j in range(1, num_classes):
        for i in range(num_images):
            a, b = dev[j][i], hst[j][i]
            assert a.shape == b.shape and np.array_equal(a[:, :4], b[:, :4])
            assert np.abs(a[:, 4].view(np.int32).astype(np.int64) - b[:, 4].view(np.int32).astype(np.int64)).max(initial=0) <= 1
            n += len(a)
    assert 0 < n <= 40 * num_images
