"""CPU suite: the oracle's restatement of the F-matrix geometric filter
(hulo::geometricMatch -> OpenMVG 1.1 GeometricFilter_FMatrix_AC, MatchUtils.cpp:372-420).
Parity against the reference is UNPINNED (no OpenMVG here, no reference test); pinned are the
7-point solution sets (OpenCV golden vectors), the epipolar error and the NFA (independent
numpy restatements) and the recovery of planted two-view geometry."""
import math
import os

import numpy as np
import pytest

from sfmlocalization_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLT_EPS = float(np.finfo(np.float32).eps)


@pytest.fixture(scope="module")
def fgold():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "fmatrix_golden.npz")))


def unit(F):
    return F / np.linalg.norm(F)


def test_seven_point_matches_opencv_solution_sets(orc, fgold):
    for x1, x2, sols, n in zip(fgold["x1"], fgold["x2"], fgold["solutions"], fgold["n_solutions"]):
        mine = [unit(F) for F in orc.seven_point(x1, x2)]
        assert len(mine) == n
        for s in sols[:n]:
            # cv2's own solutions satisfy x2^T F x1 = 0 only to ~1e-8 (ours: 1e-16), hence 1e-4
            assert min(min(np.abs(m - s).max(), np.abs(m + s).max()) for m in mine) < 1e-4


def test_seven_point_properties(orc, fgold):
    for x1, x2 in zip(fgold["x1"], fgold["x2"]):
        for F in orc.seven_point(x1, x2):
            r = [np.r_[b, 1] @ F @ np.r_[a, 1] for a, b in zip(x1, x2)]
            assert np.abs(r).max() < 1e-12 * max(1.0, np.abs(F).max())
            assert abs(np.linalg.det(unit(F))) < 1e-12


def test_seven_point_degenerate_sample(orc):
    x = np.tile(np.array([[0.1, 0.2]]), (7, 1))          # seven copies of one point: rank deficient
    assert len(orc.seven_point(x, x)) == 0


def test_precondition(orc):
    T = orc.precondition(1920, 1080)
    s = 1 / math.sqrt(1920 * 1080)
    assert np.allclose(T, [[s, 0, -960 * s], [0, s, -540 * s], [0, 0, 1]])


def test_epipolar_error_is_point_to_line_distance(orc):
    rng = np.random.default_rng(3)
    F = rng.normal(size=(3, 3))
    x1, x2 = rng.normal(size=(50, 2)), rng.normal(size=(50, 2))
    l = np.c_[x1, np.ones(50)] @ F.T
    want = (np.sum(l * np.c_[x2, np.ones(50)], axis=1)) ** 2 / (l[:, 0] ** 2 + l[:, 1] ** 2)
    assert np.allclose(orc.epipolar_errors(F, x1, x2), want, rtol=1e-12)


def np_fmatrix_nfa(e_sorted, logalpha0, max_thr=math.inf):
    N = len(e_sorted)
    loge0 = math.log10(3 * (N - 7))
    best, bk = math.inf, 7
    for k in range(8, N + 1):
        if e_sorted[k - 1] > max_thr:
            break
        logalpha = logalpha0 + 0.5 * math.log10(e_sorted[k - 1] + FLT_EPS)
        lcn = np.float32(math.log10(math.comb(N, k))) if 0 < k < N else np.float32(0)
        lck = np.float32(math.log10(math.comb(k, 7))) if k > 7 else np.float32(0)
        nfa = loge0 + logalpha * (k - 7) + float(lcn) + float(lck)
        if nfa < best:
            best, bk = nfa, k
    return best, bk


@pytest.mark.parametrize("N,thr_px", [(40, math.inf), (150, 4.0), (400, 2.0)])
def test_fmatrix_score_matches_numpy(orc, N, thr_px):
    tv = synth.two_view_matches(N, 40 + N, outlier_frac=0.4)
    w, h = tv["size"]
    T = orc.precondition(w, h)
    x1 = tv["xI"] * T[0, 0] + T[:2, 2]
    x2 = tv["xJ"] * T[0, 0] + T[:2, 2]
    logalpha0 = math.log10(2.0 * math.hypot(w, h) / (w * h) / T[0, 0])
    max_thr = math.inf if math.isinf(thr_px) else (thr_px * T[0, 0]) ** 2
    Tin = np.linalg.inv(T)
    Fn = Tin.T @ tv["F_true"] @ Tin                      # pixel F -> normalised coordinates
    e = np.sort(orc.epipolar_errors(Fn, x1, x2))
    want, wk = np_fmatrix_nfa(e, logalpha0, max_thr)
    nfa, kb, ek = orc.fmatrix_score(Fn, x1, x2, logalpha0, max_thr)
    assert kb == wk and abs(nfa - want) < 1e-9 * max(1.0, abs(want)) and ek == e[kb - 1]
    assert nfa < 0 and kb >= 0.8 * tv["inlier_mask"].sum()


@pytest.mark.parametrize("N,out", [(60, 0.3), (300, 0.5), (1000, 0.7)])
def test_acransac_recovers_planted_geometry(orc, N, out):
    tv = synth.two_view_matches(N, 7 + N, outlier_frac=out)
    r = orc.fmatrix_acransac(tv["xI"], tv["xJ"], tv["size"], tv["size"], 4.0, 1024, 5)
    assert r["ok"] and r["nfa"] < 0 and r["error_max"] <= 4.0 + 1e-9
    truth = np.flatnonzero(tv["inlier_mask"])
    assert np.isin(r["inliers"], truth).mean() > 0.93
    assert len(np.intersect1d(r["inliers"], truth)) > 0.85 * len(truth)
    # the estimated F explains the true inliers: point-to-line distance of a few pixels
    e = np.sqrt(orc.epipolar_errors(r["F"], tv["xI"][truth], tv["xJ"][truth]))
    assert np.median(e) < 1.5


def test_acransac_edge_cases(orc):
    tv = synth.two_view_matches(7, 1, outlier_frac=0.0)
    r = orc.fmatrix_acransac(tv["xI"], tv["xJ"], tv["size"], tv["size"], 4.0, 200, 1)
    assert not r["ok"] and len(r["inliers"]) == 0            # N <= 7: nothing to do
    tv = synth.two_view_matches(200, 2, outlier_frac=1.0)    # pure noise: no meaningful model
    r = orc.fmatrix_acransac(tv["xI"], tv["xJ"], tv["size"], tv["size"], 4.0, 200, 1)
    assert not r["ok"] and len(r["inliers"]) == 0 and r["nfa"] >= 0
    tv = synth.two_view_matches(200, 3, outlier_frac=0.3)
    a = orc.fmatrix_acransac(tv["xI"], tv["xJ"], tv["size"], tv["size"], 4.0, 25, 9)
    b = orc.fmatrix_acransac(tv["xI"], tv["xJ"], tv["size"], tv["size"], 4.0, 25, 9)
    assert a["ok"] == b["ok"] and np.array_equal(a["inliers"], b["inliers"]) and np.array_equal(a["F"], b["F"])
