"""GPU parity suite for K4 (hulo_ransac_transform3d: the model-merge RANSAC of mergeSfM.py) against
the oracle, replaying the reference's own sample sequence: identical winning inlier set, M within
1e-9 (the refit is least squares over the same inliers in different arithmetic)."""
import os

import numpy as np
import pytest

from tests.test_oracle_merge import similarity_case

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_affine_equals_the_reference_function_on_its_own_samples(gpu):
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "merge_golden.npz")))
    for k in range(4):
        M, inl = gpu.ransac_transform3d(g["A%d" % k], g["B%d" % k], g["par%d" % k][0], 0, g["par%d" % k][1],
                                        samples=g["samples%d" % k])
        assert np.array_equal(inl, g["inl%d" % k]), k
        assert np.abs(M - g["M%d" % k]).max() < 1e-9 * max(1.0, np.abs(g["M%d" % k]).max()), k


def test_similarity_equals_oracle(gpu, orc):
    for seed, n, of in [(5, 150, 0.5), (6, 40, 0.3), (7, 900, 0.7)]:
        A, B, M, inl, samples = similarity_case(seed, n, of, 0.01)
        Mo, io = orc.ransac_transform3d(A, B, 0.06, samples, 1.75, similarity=True)
        Mg, ig = gpu.ransac_transform3d(A, B, 0.06, 0, 1.75, similarity=True, samples=samples)
        assert np.array_equal(ig, io)
        assert np.abs(Mg - Mo).max() < 1e-9 and np.abs(Mg - M).max() < 0.02


def test_condition_number_gate_and_empty_results(gpu, orc):
    A, B, M, inl, samples = similarity_case(8, 120, 0.4, 0.01)
    # an anisotropic map: its affine fits have singular value ratio 3 > 1.75 and must be refused
    A2 = A.copy(); A2[0] *= 3.0
    # (only samples of four stray points give a conditioned fit: a handful of inliers at best)
    Mo, io = orc.ransac_transform3d(A2, B, 0.06, samples, 1.75)
    Mg, ig = gpu.ransac_transform3d(A2, B, 0.06, 0, 1.75, samples=samples)
    assert len(io) < 10 and np.array_equal(ig, io)
    if io.size:
        assert np.abs(Mg - Mo).max() < 1e-9 * max(1.0, np.abs(Mo).max())
    Mo, io = orc.ransac_transform3d(A2, B, 0.06, samples, 10.0)
    Mg, ig = gpu.ransac_transform3d(A2, B, 0.06, 0, 10.0, samples=samples)
    assert len(io) > 50 and np.array_equal(ig, io) and np.abs(Mg - Mo).max() < 1e-9
    # fewer than four points, no rounds
    assert gpu.ransac_transform3d(A[:, :3], B[:, :3], 0.06, 100)[0].size == 0
    assert gpu.ransac_transform3d(A, B, 0.06, 0)[0].size == 0


def test_own_sampler_and_python_entry_points(gpu):
    """Without a sample list the rounds are drawn on the device; the reference-named functions find
    the planted transform (ransacRound = 100 x #matches like mergeSfM.py:577)."""
    from sfmlocalization_b200 import gpu as api
    A, B, M, inl, _ = similarity_case(9, 300, 0.6, 0.01)
    for fn in (api.ransacAffineTransform, api.ransacSimilarityTransform, api.ransacTransform):
        Mh, got = fn(A, B, 0.06, 300 * 100, 1.75)
        assert np.abs(Mh - M).max() < 0.02
        assert np.isin(got, np.flatnonzero(inl)).mean() > 0.98 and len(got) > 0.9 * inl.sum()
    a = gpu.ransac_transform3d(A, B, 0.06, 5000, 1.75, seed=3)
    b = gpu.ransac_transform3d(A, B, 0.06, 5000, 1.75, seed=3)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])
