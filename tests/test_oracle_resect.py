"""CPU suite: the oracle's resection restatement (OpenMVG 1.1 AC-RANSAC + P3P).
Parity against the reference is UNPINNED (no OpenMVG here, no reference tests); these checks
pin the pieces that have independent ground truth: the P3P solution set (OpenCV golden
vectors), the NFA formula (independent numpy restatement), log-binomials (math.comb), KRt
decomposition (round trip) and end-to-end pose recovery on synthetic scenes."""
import math
import os

import numpy as np
import pytest

from sfmlocalization_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLT_EPS = float(np.finfo(np.float32).eps)


@pytest.fixture(scope="module")
def rgold():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "resect_golden.npz")))


def bearings(x2d, K, orc):
    xn = orc.normalize_points(x2d, K)
    f = np.c_[xn, np.ones(len(xn))]
    return f / np.linalg.norm(f, axis=1, keepdims=True)


def test_p3p_matches_opencv_solution_sets(orc, rgold):
    K = rgold["K"]
    for x, X, sols, n in zip(rgold["x2d"], rgold["X3d"], rgold["solutions"], rgold["n_solutions"]):
        mine = orc.p3p(bearings(x, K, orc), X)
        assert len(mine) >= n
        for s in sols[:n]:
            assert min(np.abs(m - s).max() for m in mine) < 1e-6


def test_p3p_properties(orc, rgold):
    K = rgold["K"]
    for x, X, truth in zip(rgold["x2d"], rgold["X3d"], rgold["truth"]):
        f = bearings(x, K, orc)
        mine = orc.p3p(f, X)
        assert min(np.abs(m - truth).max() for m in mine) < 1e-7          # contains the true pose
        for m in mine:
            R = m[:, :3]
            if abs(np.linalg.det(R) - 1) > 1e-6:
                continue       # real part of a complex root: not a pose, scores badly in RANSAC
            assert np.allclose(R.T @ R, np.eye(3), atol=1e-8)


def test_p3p_collinear_points_rejected(orc):
    X = np.array([[0, 0, 5.0], [1, 1, 6.0], [2, 2, 7.0]])
    f = np.array([[0, 0, 1.0], [0.1, 0.1, 1.0], [0.2, 0.2, 1.0]])
    f /= np.linalg.norm(f, axis=1, keepdims=True)
    assert len(orc.p3p(f, X)) == 0


def test_logcombi(orc):
    for n in (4, 10, 57, 300):
        for k in range(0, n + 1):
            want = 0.0 if (k >= n or k <= 0) else math.log10(math.comb(n, k))
            assert abs(orc.logcombi(k, n) - want) < 1e-4 * max(1.0, want)


def np_best_nfa(e_sorted):
    """Independent restatement of bestNFA + the ACRANSAC constants for P3P resection."""
    N = len(e_sorted)
    loge0 = math.log10(4 * (N - 3))
    best, bk = math.inf, 3
    for k in range(4, N + 1):
        logalpha = math.log10(math.pi) + math.log10(e_sorted[k - 1] + FLT_EPS)
        lcn = np.float32(math.log10(math.comb(N, k))) if 0 < k < N else np.float32(0)
        lck = np.float32(math.log10(math.comb(k, 3))) if k > 3 else np.float32(0)
        nfa = loge0 + logalpha * (k - 3) + float(lcn) + float(lck)
        if nfa < best:
            best, bk = nfa, k
    return best, bk


def test_nfa_known_answer(orc):
    # ten residuals: six tight inliers then a jump
    e = np.array([1e-8, 2e-8, 2.5e-8, 4e-8, 5e-8, 9e-8, 3e-3, 8e-3, 2e-2, 9e-2])
    nfa, k = orc.best_nfa(e)
    want, wk = np_best_nfa(e)
    assert k == wk == 6
    assert abs(nfa - want) < 1e-5
    assert nfa < 0


def test_nfa_random_lists(orc):
    rng = np.random.default_rng(5)
    for N in (12, 60, 333):
        e = np.sort(np.concatenate([rng.uniform(0, 1e-6, N // 2), rng.uniform(0, 0.3, N - N // 2)]))
        nfa, k = orc.best_nfa(e)
        want, wk = np_best_nfa(e)
        assert k == wk and abs(nfa - want) < 1e-4 * max(1.0, abs(want))


def test_residuals_and_scoring(orc):
    sc = synth.resection_scene(200, 3, outlier_frac=0.4)
    K = sc["K"]
    xn = orc.normalize_points(sc["x2d"], K)
    M = np.c_[sc["R"], sc["t"].reshape(3, 1)]
    e = orc.residuals(M, xn, sc["X3d"])
    px = np.sqrt(e) * K[0, 0]
    assert np.median(px[sc["inlier_mask"]]) < 2.0 and np.median(px[~sc["inlier_mask"]]) > 50
    nfa, kb, ek, ni = orc.score_hypotheses(M[None], xn, sc["X3d"], thr2=(4.0 / K[0, 0]) ** 2)
    want, wk = np_best_nfa(np.sort(e))
    assert kb[0] == wk and abs(nfa[0] - want) < 1e-6 * abs(want)
    assert abs(int(kb[0]) - int(sc["inlier_mask"].sum())) <= 6
    assert ni[0] == int((e <= (4.0 / K[0, 0]) ** 2).sum())


@pytest.mark.parametrize("N,outl", [(100, 0.3), (500, 0.5), (500, 0.7)])
def test_acransac_recovers_pose(orc, N, outl):
    sc = synth.resection_scene(N, 40 + N, outlier_frac=outl)
    r = orc.acransac(sc["x2d"], sc["X3d"], sc["K"], max_iter=4096, seed=7)
    assert r["ok"] and r["nfa"] < 0
    K, R, t, c = orc.krt_from_p(r["P"])
    c_true = -sc["R"].T @ sc["t"]
    assert np.linalg.norm(c - c_true) < 0.05
    assert np.abs(R - sc["R"]).max() < 5e-3
    assert np.allclose(K, sc["K"], atol=1e-6)
    inl = set(r["inliers"].tolist())
    truth = set(np.nonzero(sc["inlier_mask"])[0].tolist())
    assert len(inl & truth) >= 0.85 * len(truth)
    assert r["error_max"] < 5.0


def test_acransac_too_few_points(orc):
    sc = synth.resection_scene(3, 1, outlier_frac=0.0)
    r = orc.acransac(sc["x2d"], sc["X3d"], sc["K"])
    assert not r["ok"] and len(r["inliers"]) == 0


def test_acransac_pure_outliers_not_meaningful(orc):
    sc = synth.resection_scene(60, 9, outlier_frac=1.0)
    r = orc.acransac(sc["x2d"], sc["X3d"], sc["K"], max_iter=512, seed=3)
    assert not r["ok"]


def test_krt_round_trip(orc):
    sc = synth.resection_scene(4, 2)
    P = sc["K"] @ np.c_[sc["R"], sc["t"].reshape(3, 1)]
    for scale in (1.0, -2.5):
        K, R, t, c = orc.krt_from_p(scale * P)
        assert np.allclose(K, sc["K"], atol=1e-8) and np.allclose(R, sc["R"], atol=1e-10)
        assert np.allclose(t, sc["t"], atol=1e-10) and np.allclose(c, -sc["R"].T @ sc["t"], atol=1e-10)
