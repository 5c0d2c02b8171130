"""GPU parity suite for K3 (F-matrix geometric filter, hulo::geometricMatch) through the
C-ABI against the CPU oracle.  The device sampler is the oracle's counter-based stream and both
keep the narrowed pool in index order, so for a given seed both run the same schedule; the fp64
solver differs by libm-vs-device rounding and the residual keys are fp32, so a near-tie (two
models or two k of almost equal NFA, a residual on the precision bound) can send the traces
apart.  Hence, per pair: valid flag identical; inlier SETS identical for nearly all pairs (the
order of the seven zero-residual sample points inside the list is rounding noise on both sides)
and overlapping otherwise; when the sets agree: log10 NFA within 2e-3 * max(1, |nfa|), error_max
within 1e-3 px, F equal up to scale to 1e-6; error_max never above the bound; the list is in
ascending residual order."""
import math

import numpy as np
import pytest

from sfmlocalization_b200 import synth

pytestmark = pytest.mark.gpu


def batch(specs, seed0):
    xs, ys, off, truth = [], [], [0], []
    for k, (N, outl) in enumerate(specs):
        tv = synth.two_view_matches(N, seed0 + k, outlier_frac=outl) if N else dict(
            xI=np.zeros((0, 2)), xJ=np.zeros((0, 2)), inlier_mask=np.zeros(0, bool))
        xs.append(tv["xI"]); ys.append(tv["xJ"]); off.append(off[-1] + N); truth.append(tv["inlier_mask"])
    return np.concatenate(xs), np.concatenate(ys), np.array(off, np.uint64), truth


def unit(F):
    n = np.linalg.norm(F)
    return F / n if n > 0 else F


def compare(gpu, orc, specs, precision, max_iter, seed, seed0=100):
    xI, xJ, off, truth = batch(specs, seed0)
    w, h = synth.IMAGE_WH
    sizes = np.tile(np.array([w, h, w, h], np.int32), (len(specs), 1))
    r = gpu.geometric_filter(xI, xJ, off, sizes, precision, max_iter, seed)
    exact = 0
    for p in range(len(specs)):
        a, b = int(off[p]), int(off[p + 1])
        o = orc.fmatrix_acransac(xI[a:b], xJ[a:b], (w, h), (w, h), precision, max_iter, seed + 1000003 * p)
        assert bool(r["valid"][p]) == o["ok"], (p, specs[p])
        gi, oi = r["inliers"][p], o["inliers"]
        if len(oi) == 0:
            assert len(gi) == 0
            continue
        union = len(np.union1d(gi, oi))
        assert len(np.intersect1d(gi, oi)) >= 0.6 * union, (p, specs[p], len(gi), len(oi))
        if not math.isinf(precision):
            assert r["error_max"][p] <= precision * (1 + 1e-6)
        # the list is in ascending residual order under the model returned with it
        e = np.sqrt(orc.epipolar_errors(r["F"][p], xI[a:b], xJ[a:b]))[gi]
        assert (np.diff(e) >= -1e-4).all() and abs(e[-1] - r["error_max"][p]) <= 1e-3
        if np.array_equal(np.sort(gi), np.sort(oi)):
            exact += 1
            assert abs(r["nfa"][p] - o["nfa"]) <= 2e-3 * max(1.0, abs(o["nfa"]))
            assert abs(r["error_max"][p] - o["error_max"]) <= 1e-3
            Fg, Fo = unit(r["F"][p]), unit(o["F"])
            assert min(np.abs(Fg - Fo).max(), np.abs(Fg + Fo).max()) < 1e-6
            # beyond the zero-residual sample points the two lists are the same sequence
            assert np.array_equal(gi[e > 1e-6], oi[np.isin(oi, gi[e > 1e-6])])
    return exact, r


def test_batch_matches_oracle_localization_setting(gpu, orc):
    """ransacRound 25, geomPrec 4 px (LocalizeParam.py:35): pairs of every size a query produces."""
    specs = [(16, 0.2), (17, 0.0), (24, 0.3), (40, 0.5), (64, 0.4), (100, 0.3), (129, 0.6), (200, 0.5),
             (333, 0.4), (512, 0.5), (700, 0.7), (1000, 0.6), (30, 1.0), (300, 1.0), (18, 0.1), (1500, 0.5)]
    exact, r = compare(gpu, orc, specs, 4.0, 25, 7)
    found = int((r["n_inliers"] > 0).sum())
    assert found >= 9 and exact >= found - 1
    assert r["valid"].sum() >= 7


def test_batch_matches_oracle_reconstruction_setting(gpu, orc):
    """ransacRound 500 (ReconstructParam.py:76) on image-pair sized match lists."""
    specs = [(60, 0.3), (250, 0.5), (900, 0.6), (2000, 0.5), (3000, 0.8)]
    exact, r = compare(gpu, orc, specs, 4.0, 500, 11, seed0=300)
    assert r["valid"].all() and exact >= 4


def test_no_precision_bound(gpu, orc):
    specs = [(50, 0.3), (400, 0.5), (1024, 0.4)]
    exact, r = compare(gpu, orc, specs, math.inf, 200, 3, seed0=500)
    assert r["valid"].all() and exact >= 2


@pytest.mark.parametrize("max_iter", [1, 9, 10, 11, 25])
def test_iteration_budget_edges(gpu, orc, max_iter):
    specs = [(80, 0.2), (80, 0.6), (200, 0.9)]
    compare(gpu, orc, specs, 4.0, max_iter, 5, seed0=700)


def test_degenerate_pairs(gpu, orc):
    """Empty pairs, pairs at and below the minimal sample, identical points."""
    specs = [(0, 0.0), (5, 0.0), (7, 0.0), (8, 0.0), (0, 0.0), (120, 0.3)]
    exact, r = compare(gpu, orc, specs, 4.0, 25, 2, seed0=800)
    assert not r["valid"][:3].any() and r["n_inliers"][:3].sum() == 0 and r["valid"][5]
    # all matches at the same pixel: every sample is rank deficient, nothing is found, nothing hangs
    x = np.tile(np.array([[100.0, 200.0]]), (40, 1))
    w, h = synth.IMAGE_WH
    r = gpu.geometric_filter(x, x, np.array([0, 40], np.uint64), np.array([[w, h, w, h]], np.int32), 4.0, 25, 1)
    o = orc.fmatrix_acransac(x, x, (w, h), (w, h), 4.0, 25, 1)
    assert not r["valid"][0] and not o["ok"]
    # no pairs at all
    r = gpu.geometric_filter(np.zeros((0, 2)), np.zeros((0, 2)), np.array([0], np.uint64), np.zeros((0, 4), np.int32))
    assert len(r["valid"]) == 0


def test_rejects_bad_arguments(gpu):
    from sfmlocalization_b200.gpu import HuloError
    x = np.zeros((20, 2))
    with pytest.raises(HuloError):
        gpu.geometric_filter(x, x, np.array([0, 20], np.uint64), np.array([[0, 10, 10, 10]], np.int32))
    with pytest.raises(HuloError):
        gpu.geometric_filter(x, x, np.array([0, 20], np.uint64), np.array([[10, 10, 10, 10]], np.int32), -1.0)
    big = np.zeros((16385, 2))
    with pytest.raises(HuloError):
        gpu.geometric_filter(big, big, np.array([0, 16385], np.uint64), np.array([[10, 10, 10, 10]], np.int32))


def test_different_image_sizes_and_many_pairs(gpu, orc):
    """300 pairs in one launch (more blocks than SMs), two image sizes."""
    rng = np.random.default_rng(4)
    specs = [(int(rng.integers(16, 260)), float(rng.choice([0.2, 0.5, 0.8, 1.0]))) for _ in range(300)]
    xI, xJ, off, truth = batch(specs, 2000)
    w, h = synth.IMAGE_WH
    sizes = np.tile(np.array([w, h, w, h], np.int32), (300, 1))
    sizes[::2] = [w, h, 2 * w, 2 * h]          # image J declared larger: changes s2, logalpha0, the bound
    r = gpu.geometric_filter(xI, xJ, off, sizes, 4.0, 25, 9)
    agree = 0
    for p in range(300):
        a, b = int(off[p]), int(off[p + 1])
        o = orc.fmatrix_acransac(xI[a:b], xJ[a:b], tuple(sizes[p, :2]), tuple(sizes[p, 2:]), 4.0, 25, 9 + 1000003 * p)
        assert bool(r["valid"][p]) == o["ok"], p
        agree += int(np.array_equal(np.sort(r["inliers"][p]), np.sort(o["inliers"])))
        if o["ok"]:
            t = np.flatnonzero(truth[p])
            assert np.isin(r["inliers"][p], t).mean() > 0.85
    assert agree >= 285


def test_large_pair_uses_the_big_shared_memory_buffer(gpu, orc):
    """9000 putative matches in one pair: 128 KB of sort keys (dynamic shared memory opt-in)."""
    specs = [(9000, 0.5), (40, 0.3)]
    exact, r = compare(gpu, orc, specs, 4.0, 200, 13, seed0=900)
    assert r["valid"][0] and r["n_inliers"][0] > 3000 and exact >= 1


def test_inliers_agree_with_opencv_anchor(gpu):
    """Independent anchor (tests/golden/make_golden_anchors.py): the inlier sets of the device F-matrix
    AC-RANSAC against cv2.findFundamentalMat(FM_RANSAC)'s at the threshold the fp64 restatement
    estimated, and against the planted truth."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    a = dict(np.load(os.path.join(root, "tests", "golden", "anchors_golden.npz")))
    w, h = synth.IMAGE_WH
    for k, (N, outl, seed) in enumerate(a["fmat_cases"]):
        sc = synth.two_view_matches(int(N), int(seed), outlier_frac=float(outl))
        r = gpu.geometric_filter(sc["xI"], sc["xJ"], np.array([0, int(N)], np.uint64),
                                 np.array([[w, h, w, h]], np.int32), 16.0, 1024, 1)
        assert r["valid"][0]
        mine = np.zeros(int(N), bool); mine[r["inliers"][0]] = True
        cv = a["fmat_%d_cv2_inliers" % k]
        assert abs(r["error_max"][0] - float(a["fmat_%d_threshold_px" % k])) < 0.75
        assert (cv & mine).sum() / cv.sum() >= 0.93
        assert (cv & mine).sum() / (cv | mine).sum() >= 0.85
        assert (mine & sc["inlier_mask"]).sum() / sc["inlier_mask"].sum() >= 0.90
