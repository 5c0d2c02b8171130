"""CPU suite: independent anchors for the restatements whose upstream source (OpenMVG 1.1) cannot be
obtained here (VERDICT r1 "parity unpinned").  Three kinds of evidence, none of them the oracle
checking itself:
  1. the a-contrario score (bestNFA, log-binomials) against an exact-integer restatement
     (math.comb, correctly rounded log10 of the exact binomial) on 1000 random residual lists with
     ties and threshold cuts;
  2. the FINAL inlier sets of the resection and F-matrix AC-RANSAC against OpenCV's own robust
     estimators at the threshold AC-RANSAC estimated (fixtures made by
     tests/golden/make_golden_anchors.py with cv2.solvePnPRansac(P3P) / cv2.findFundamentalMat);
  3. against the planted ground truth of the synthetic scenes.
What stays unpinned is stated in DESIGN.md section 5."""
import math
import os

import numpy as np
import pytest

from sfmlocalization_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLT_EPS = float(np.finfo(np.float32).eps)


@pytest.fixture(scope="module")
def anchors():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "anchors_golden.npz")))


def exact_nfa_table(e, n_sample, n_models, logalpha0, max_thr):
    """NFA_k for k = n_sample + 1 .. N from exact binomials: log10 of a Python int is correctly
    rounded, so the only fp64 rounding left is in the sum."""
    N = len(e)
    loge0 = math.log10(n_models * (N - n_sample))
    out = {}
    for k in range(n_sample + 1, N + 1):
        if e[k - 1] > max_thr:
            break
        logalpha = logalpha0 + math.log10(e[k - 1] + FLT_EPS)
        out[k] = loge0 + logalpha * (k - n_sample) + math.log10(math.comb(N, k)) + math.log10(math.comb(k, n_sample))
    return out


def test_log_binomials_against_exact_integers(orc):
    rng = np.random.default_rng(1)
    for _ in range(2000):
        n = int(rng.integers(1, 4000)); k = int(rng.integers(0, n + 1))
        want = math.log10(math.comb(n, k)) if 0 < k < n else 0.0
        assert abs(orc.logcombi(k, n) - want) <= 2e-6 * max(1.0, want)          # float32 table entries


def test_best_nfa_against_exact_restatement(orc):
    """1000 residual lists: random sizes, inlier/outlier mixtures, exact ties, threshold cuts."""
    import ctypes as C
    rng = np.random.default_rng(2)
    lib = orc.lib()
    checked_cut = checked_tie = 0
    for trial in range(1000):
        N = int(rng.integers(5, 400))
        n_in = int(rng.integers(0, N + 1))
        e = np.concatenate([rng.random(n_in) ** 2 * 1e-5, rng.random(N - n_in) * 0.2 + 1e-4])
        if trial % 3 == 0:                              # exact ties
            e = np.round(e, 5 if trial % 2 else 4)
            checked_tie += 1
        e = np.sort(e)
        max_thr = float(np.inf) if trial % 4 else float(e[int(rng.integers(0, N))])
        checked_cut += int(np.isfinite(max_thr))
        lcn = np.empty(N + 1, np.float32); lck = np.empty(N + 1, np.float32)
        lib.orc_make_logcombi(C.c_size_t(N), lcn.ctypes.data_as(C.c_void_p), lck.ctypes.data_as(C.c_void_p))
        kb = C.c_size_t(0)
        logalpha0 = math.log10(math.pi)
        got = lib.orc_best_nfa(e.ctypes.data_as(C.c_void_p), C.c_size_t(N), C.c_double(logalpha0),
                               C.c_double(math.log10(4.0 * (N - 3))), C.c_double(max_thr),
                               lcn.ctypes.data_as(C.c_void_p), lck.ctypes.data_as(C.c_void_p), C.c_double(1.0),
                               C.byref(kb))
        table = exact_nfa_table(e, 3, 4, logalpha0, max_thr)
        if not table:
            assert math.isinf(got) and kb.value == 3
            continue
        best = min(table.values())
        # value: the float32 log-binomial tables carry ~1e-7 relative error each
        assert abs(got - best) <= 5e-5 + 2e-6 * abs(best)
        # arg-min: the k chosen is a minimiser up to that table error, and never beyond the cut
        assert kb.value in table
        assert table[kb.value] - best <= 1e-4 + 4e-6 * abs(best)
        # first minimum wins: no smaller k scores clearly better or equal
        assert all(table[k] > table[kb.value] - (1e-4 + 4e-6 * abs(best)) for k in table if k < kb.value)
    assert checked_cut > 100 and checked_tie > 100


def jaccard(a, b):
    return (a & b).sum() / max(1, (a | b).sum())


def test_resection_inliers_agree_with_opencv_and_truth(orc, anchors):
    for k, (N, outl, seed) in enumerate(anchors["resect_cases"]):
        sc = synth.resection_scene(int(N), int(seed), outlier_frac=float(outl))
        r = orc.acransac(sc["x2d"], sc["X3d"], sc["K"], max_iter=4096, seed=1)
        assert r["ok"]
        # the threshold is re-estimated here; the fixture was made at the same one
        assert abs(r["error_max"] - float(anchors["resect_%d_threshold_px" % k])) < 1e-9
        mine = np.zeros(int(N), bool); mine[r["inliers"]] = True
        cv = anchors["resect_%d_cv2_inliers" % k]
        # OpenCV keeps the best unrefined minimal model of its own sampling: its inlier set is a little
        # smaller and (almost) contained in ours
        assert (cv & mine).sum() / cv.sum() >= 0.97
        assert jaccard(cv, mine) >= 0.90
        assert jaccard(sc["inlier_mask"], mine) >= 0.97
        assert mine.sum() >= cv.sum() - 2


def test_fmatrix_inliers_agree_with_opencv_and_truth(orc, anchors):
    for k, (N, outl, seed) in enumerate(anchors["fmat_cases"]):
        sc = synth.two_view_matches(int(N), int(seed), outlier_frac=float(outl))
        r = orc.fmatrix_acransac(sc["xI"], sc["xJ"], sc["size"], sc["size"], 16.0, 1024, 1)
        assert r["ok"]
        assert abs(r["error_max"] - float(anchors["fmat_%d_threshold_px" % k])) < 1e-9
        mine = np.zeros(int(N), bool); mine[r["inliers"]] = True
        cv = anchors["fmat_%d_cv2_inliers" % k]
        assert (cv & mine).sum() / cv.sum() >= 0.95
        assert jaccard(cv, mine) >= 0.88
        # outliers that happen to lie near the epipolar line are legitimately accepted
        assert (mine & sc["inlier_mask"]).sum() / sc["inlier_mask"].sum() >= 0.93
