"""CPU suite: the oracle's restatement of the 3D-3D model-merge RANSAC.
The affine flavour is PINNED against the reference's own function: tests/golden/merge_golden.npz
was produced by executing ransacAffineTransform of mergeSfM.py:344-388 itself (see
make_golden_merge.py), with the 4-point samples of every round recorded."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mgold():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "merge_golden.npz")))


def test_affine_restatement_equals_the_reference_function(orc, mgold):
    for k in range(4):
        M, inl = orc.ransac_transform3d(mgold["A%d" % k], mgold["B%d" % k], mgold["par%d" % k][0],
                                        mgold["samples%d" % k], mgold["par%d" % k][1])
        assert np.array_equal(inl, mgold["inl%d" % k])
        assert np.array_equal(M, mgold["M%d" % k])               # same numpy calls in the same order


def similarity_case(seed, n, outlier_frac, noise):
    rng = np.random.default_rng(seed)
    B = rng.normal(size=(3, n)) * 4
    Q = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    Q *= np.sign(np.linalg.det(Q))
    M = np.hstack([1.7 * Q, rng.normal(size=(3, 1)) * 2])
    A = M @ np.vstack([B, np.ones((1, n))]) + rng.normal(size=(3, n)) * noise
    out = rng.random(n) < outlier_frac
    A[:, out] = rng.normal(size=(3, int(out.sum()))) * 4
    samples = np.array([rng.choice(n, 4, replace=False) for _ in range(600)])
    return A, B, M, ~out, samples


def test_superimposition_matrix_is_the_least_squares_similarity(orc):
    A, B, M, inl, _ = similarity_case(3, 50, 0.0, 0.0)
    S = orc.superimposition_matrix(B, A)
    assert np.allclose(S[:3], M, atol=1e-10) and np.allclose(S[3], [0, 0, 0, 1])
    # a reflection in the data must not produce a reflection in the result
    A2 = A.copy(); A2[0] *= -1
    S2 = orc.superimposition_matrix(B, A2)
    assert np.linalg.det(S2[:3, :3]) > 0


def test_similarity_ransac_recovers_the_transform(orc):
    A, B, M, inl, samples = similarity_case(5, 150, 0.5, 0.01)
    Mh, got = orc.ransac_transform3d(A, B, 0.06, samples, 1.75, similarity=True)
    assert np.abs(Mh - M).max() < 0.02
    assert np.isin(got, np.flatnonzero(inl)).mean() > 0.98 and len(got) > 0.9 * inl.sum()
    none = orc.ransac_transform3d(A[:, :3], B[:, :3], 0.06, np.zeros((0, 4), int))
    assert none[0].size == 0 and none[1].size == 0
