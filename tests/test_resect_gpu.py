"""GPU parity suite for K2 (resection hypothesis scoring), the device P3P solver and the
batched AC-RANSAC driver, through the C-ABI, against the fp64 CPU oracle.
Tolerance (BASELINE.json): residuals within 1e-4 px of the fp64 restatement."""
import math

import numpy as np
import pytest

from sfmlocalization_b200 import synth

pytestmark = pytest.mark.gpu
TOL_PX = 1e-4


def hypotheses(orc, sc, T, seed):
    """P3P models of T seeded triplets (oracle), flattened H x 3 x 4."""
    xn = orc.normalize_points(sc["x2d"], sc["K"])
    tri = synth.sample_triplets(len(xn), T, seed)
    models = []
    for a in tri:
        f = np.c_[xn[a], np.ones(3)]
        f /= np.linalg.norm(f, axis=1, keepdims=True)
        models += list(orc.p3p(f, sc["X3d"][a]))
    return np.array(models), tri


@pytest.mark.parametrize("N,outl", [(100, 0.3), (500, 0.5), (2000, 0.7)])
def test_residuals_within_1e4_px(gpu, orc, N, outl):
    sc = synth.resection_scene(N, 70 + N, outlier_frac=outl)
    models, _ = hypotheses(orc, sc, 64, 1)
    models = models[np.abs(models).max(axis=(1, 2)) < 1e3][:200]
    res = gpu.resection_residuals(models, sc["x2d"], sc["X3d"], sc["K"])
    xn = orc.normalize_points(sc["x2d"], sc["K"])
    fx = sc["K"][0, 0]
    worst = 0.0
    for h, M in enumerate(models):
        want = np.sqrt(orc.residuals(M, xn, sc["X3d"])) * fx
        # absolute 1e-4 px; beyond ~800 px one fp32 ulp of the value itself exceeds that
        tol = np.maximum(TOL_PX, want * 2.0 ** -22)
        ok = np.isfinite(want)
        assert (np.abs(res[h][ok] - want[ok]) <= tol[ok]).all()
        small = ok & (want < 100)
        if small.any():
            worst = max(worst, float(np.abs(res[h][small] - want[small]).max()))
    assert worst <= TOL_PX


@pytest.mark.parametrize("N,outl", [(12, 0.2), (100, 0.3), (500, 0.5), (2000, 0.5), (2049, 0.7)])
def test_scores_vs_oracle(gpu, orc, N, outl):
    sc = synth.resection_scene(N, 90 + N, outlier_frac=outl)
    models, _ = hypotheses(orc, sc, 96, 2)
    true = np.c_[sc["R"], sc["t"].reshape(3, 1)][None]
    models = np.concatenate([true, models[np.abs(models).max(axis=(1, 2)) < 1e3]])
    thr = 4.0
    nfa, kb, ek, ni = gpu.score_resection(models, sc["x2d"], sc["X3d"], sc["K"], thr_px=thr)
    xn = orc.normalize_points(sc["x2d"], sc["K"])
    fx = sc["K"][0, 0]
    onfa, okb, oek, oni = orc.score_hypotheses(models, xn, sc["X3d"], thr2=(thr / fx) ** 2)
    assert nfa[0] < 0 and abs(int(kb[0]) - int(sc["inlier_mask"].sum())) <= max(6, N // 50)
    fin = np.isfinite(onfa)
    assert (np.isfinite(nfa) == fin).all()
    assert np.abs(nfa[fin] - onfa[fin]).max() <= 2e-3 + 2e-6 * np.abs(onfa[fin]).max()
    assert int(np.argmin(nfa)) == int(np.argmin(onfa))
    same = kb == okb
    assert same.mean() > 0.97
    assert same[0]
    oek_px = np.sqrt(oek) * fx
    assert np.abs(ek[same & fin] - oek_px[same & fin]).max() <= np.maximum(TOL_PX, 1e-6 * oek_px[same & fin]).max()
    # inlier counts at a fixed threshold: identical except points within tolerance of it
    assert np.abs(ni - oni).max() <= 1


def test_scores_degenerate_inputs(gpu):
    sc = synth.resection_scene(3, 5, outlier_frac=0.0)
    M = np.c_[sc["R"], sc["t"].reshape(3, 1)][None]
    nfa, kb, ek, ni = gpu.score_resection(M, sc["x2d"], sc["X3d"], sc["K"])
    assert np.isinf(nfa[0]) and kb[0] == 3
    bad = np.full((1, 3, 4), np.nan)
    sc = synth.resection_scene(50, 6)
    nfa, kb, ek, ni = gpu.score_resection(bad, sc["x2d"], sc["X3d"], sc["K"])
    assert np.isinf(nfa[0])


def test_device_p3p_vs_oracle(gpu, orc):
    sc = synth.resection_scene(400, 11, outlier_frac=0.3)
    tri = synth.sample_triplets(400, 256, 4)
    models, nm = gpu.p3p(tri, sc["x2d"], sc["X3d"], sc["K"])
    xn = orc.normalize_points(sc["x2d"], sc["K"])
    checked = 0
    for t, a in enumerate(tri):
        f = np.c_[xn[a], np.ones(3)]
        f /= np.linalg.norm(f, axis=1, keepdims=True)
        want = orc.p3p(f, sc["X3d"][a])
        assert nm[t] == len(want)
        for m in range(nm[t]):
            scale = max(1.0, np.abs(want[m]).max())
            if scale > 1e6:
                continue
            assert np.abs(models[t, m] - want[m]).max() <= 1e-7 * scale
            checked += 1
    assert checked > 700


def test_device_p3p_golden_triplets(gpu):
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = dict(np.load(os.path.join(root, "tests", "golden", "resect_golden.npz")))
    for x, X, sols, n in zip(g["x2d"], g["X3d"], g["solutions"], g["n_solutions"]):
        models, nm = gpu.p3p(np.array([[0, 1, 2]]), x, X, g["K"])
        for s in sols[:n]:
            assert min(np.abs(models[0, m] - s).max() for m in range(nm[0])) < 1e-6


@pytest.mark.parametrize("N,outl,seed", [(100, 0.3, 1), (500, 0.5, 2), (500, 0.7, 3), (2000, 0.7, 4)])
def test_acransac_recovers_pose(gpu, orc, N, outl, seed):
    sc = synth.resection_scene(N, 300 + N + seed, outlier_frac=outl)
    r = gpu.resect_acransac(sc["x2d"], sc["X3d"], sc["K"], max_iter=4096, seed=seed)
    assert r["found"]
    K, R, t, c = orc.krt_from_p(r["P"])
    c_true = -sc["R"].T @ sc["t"]
    assert np.linalg.norm(c - c_true) < 0.05
    assert np.abs(R - sc["R"]).max() < 5e-3
    inl = set(r["inliers"].tolist())
    truth = set(np.nonzero(sc["inlier_mask"])[0].tolist())
    assert len(inl & truth) >= 0.85 * len(truth)
    assert len(inl - truth) <= max(3, 0.05 * len(truth))
    # same model quality as the sequential CPU restatement on the same data
    o = orc.acransac(sc["x2d"], sc["X3d"], sc["K"], max_iter=4096, seed=seed)
    assert o["ok"]
    assert abs(len(r["inliers"]) - len(o["inliers"])) <= max(5, 0.05 * len(o["inliers"]))
    assert r["error_max"] < 2.0 * o["error_max"] + 0.5


def test_acransac_above_the_register_sort_limit(gpu, orc):
    """More than 4096 correspondences: the host-drawn three-kernel waves (shared-memory sort)
    instead of the two-launch waves of the smaller problems."""
    sc = synth.resection_scene(5000, 977, outlier_frac=0.5)
    r = gpu.resect_acransac(sc["x2d"], sc["X3d"], sc["K"], max_iter=1024, seed=3)
    assert r["found"]
    _, R, _, c = orc.krt_from_p(r["P"])
    assert np.linalg.norm(c - (-sc["R"].T @ sc["t"])) < 0.05
    assert np.abs(R - sc["R"]).max() < 5e-3
    inl = set(r["inliers"].tolist())
    truth = set(np.nonzero(sc["inlier_mask"])[0].tolist())
    assert len(inl & truth) >= 0.85 * len(truth)
    # and a problem just inside the limit gives the same answer through the batch entry point,
    # whose waves are host-drawn: same draws, same models, same decisions
    sc = synth.resection_scene(4096, 978, outlier_frac=0.5)
    one = gpu.resect_acransac(sc["x2d"], sc["X3d"], sc["K"], max_iter=1024, seed=4)
    many = gpu.resect_acransac_batch(sc["x2d"], sc["X3d"], [0, 4096], sc["K"], max_iter=1024, seeds=[4])[0]
    assert one["found"] and many["found"]
    assert np.array_equal(one["P"], many["P"]) and np.array_equal(one["inliers"], many["inliers"])


def test_acransac_not_found_cases(gpu):
    sc = synth.resection_scene(3, 1, outlier_frac=0.0)
    r = gpu.resect_acransac(sc["x2d"], sc["X3d"], sc["K"])
    assert not r["found"] and len(r["inliers"]) == 0
    sc = synth.resection_scene(60, 9, outlier_frac=1.0)
    r = gpu.resect_acransac(sc["x2d"], sc["X3d"], sc["K"], max_iter=512, seed=3)
    assert not r["found"]


@pytest.mark.parametrize("N,outl,max_iter,seed", [(60, 0.3, 4096, 1), (300, 0.5, 4096, 2), (1000, 0.6, 1024, 3),
                                                  (2000, 0.7, 4096, 4), (12, 0.2, 200, 5), (500, 0.0, 64, 6)])
def test_sequential_schedule_runs_the_oracle_trace(gpu, orc, N, outl, max_iter, seed):
    """hulo_resect_acransac_sequential keeps ACRANSAC's schedule to the letter, and both sides keep
    the narrowed pool in index order, so for one seed the device commits the draws the sequential
    CPU restatement commits: same model (P to 1e-9 relative), same inlier set; the order of the list
    agrees beyond the three zero-residual sample points."""
    sc = synth.resection_scene(N, 300 + N, outlier_frac=outl)
    g = gpu.resect_acransac(sc["x2d"], sc["X3d"], sc["K"], max_iter=max_iter, seed=seed, sequential=True)
    o = orc.acransac(sc["x2d"], sc["X3d"], sc["K"], max_iter=max_iter, seed=seed)
    assert g["found"] == (o["ok"] and len(o["inliers"]) > 7.5)
    assert np.array_equal(np.sort(g["inliers"]), np.sort(o["inliers"]))
    if len(o["inliers"]):
        assert np.abs(g["P"] - o["P"]).max() <= 1e-9 * np.abs(o["P"]).max()
        assert abs(g["error_max"] - o["error_max"]) <= 1e-3
        assert np.array_equal(g["inliers"][3:], o["inliers"][3:]) or len(np.setdiff1d(g["inliers"][:6], o["inliers"][:6])) <= 3


def test_sequential_schedule_degenerate_inputs(gpu, orc):
    sc = synth.resection_scene(3, 1, outlier_frac=0.0)
    g = gpu.resect_acransac(sc["x2d"], sc["X3d"], sc["K"], sequential=True)
    assert not g["found"] and len(g["inliers"]) == 0
    sc = synth.resection_scene(200, 2, outlier_frac=1.0)
    g = gpu.resect_acransac(sc["x2d"], sc["X3d"], sc["K"], max_iter=512, seed=3, sequential=True)
    o = orc.acransac(sc["x2d"], sc["X3d"], sc["K"], max_iter=512, seed=3)
    assert not g["found"] and len(g["inliers"]) == len(o["inliers"]) == 0


# ---------------------------------------------------------------- many resections in one call
def _batch_problems(specs, seed0=500):
    scs = [synth.resection_scene(N, seed0 + i, outlier_frac=o) for i, (N, o) in enumerate(specs)]
    off = np.zeros(len(scs) + 1, np.uint64)
    off[1:] = np.cumsum([len(s["x2d"]) for s in scs])
    x2d = np.concatenate([s["x2d"] for s in scs]) if scs else np.zeros((0, 2))
    X3d = np.concatenate([s["X3d"] for s in scs]) if scs else np.zeros((0, 3))
    return scs, off, x2d, X3d


def test_batch_resection_equals_single_calls(gpu):
    """hulo_resect_acransac_batch runs every problem on the schedule of hulo_resect_acransac:
    identical bits, whatever phases the problems of a wave are in.  The mix: clean sets (leave the
    global phase in wave 1), outlier-heavy ones (several growing waves), pure outliers (never a
    model: the reserved iterations stay global), every size class of the scoring kernel, sets too
    small to resect, an empty one, and one above the register-sort limit (runs alone)."""
    specs = [(100, 0.3), (700, 0.5), (2000, 0.7), (40, 0.0), (300, 1.0), (3, 0.0), (0, 0.0), (1500, 0.85),
             (256, 0.2), (257, 0.6), (4096, 0.5), (4500, 0.4), (12, 0.5), (600, 0.9), (1024, 0.1), (5, 0.0)]
    scs, off, x2d, X3d = _batch_problems(specs)
    Ks = np.stack([s["K"] * (1.0 if i % 2 == 0 else 1.0) for i, s in enumerate(scs)])
    Ks[1::2, 0, 0] *= 1.01          # per-problem intrinsics are honoured
    seeds = np.arange(len(specs), dtype=np.uint64) * 7919 + 11
    got = gpu.resect_acransac_batch(x2d, X3d, off, Ks, 4096, seeds=seeds)
    n_found = 0
    for p, sc in enumerate(scs):
        want = gpu.resect_acransac(sc["x2d"], sc["X3d"], Ks[p], 4096, int(seeds[p]))
        assert got[p]["found"] == want["found"], p
        assert np.array_equal(got[p]["inliers"], want["inliers"]), p
        assert got[p]["error_max"] == want["error_max"], p
        if want["found"]:
            assert np.array_equal(got[p]["P"], want["P"]), p
            n_found += 1
    assert n_found >= 9


def test_batch_resection_default_seeds_and_small_budgets(gpu):
    scs, off, x2d, X3d = _batch_problems([(200, 0.4)] * 5 + [(900, 0.6)] * 3, seed0=900)
    for max_iter in (1, 9, 64, 1000):
        got = gpu.resect_acransac_batch(x2d, X3d, off, scs[0]["K"], max_iter, seed=42)
        for p, sc in enumerate(scs):
            want = gpu.resect_acransac(sc["x2d"], sc["X3d"], sc["K"], max_iter, 42 + 1000003 * p)
            assert got[p]["found"] == want["found"]
            assert np.array_equal(got[p]["inliers"], want["inliers"])
            if want["found"]:
                assert np.array_equal(got[p]["P"], want["P"])
    assert gpu.resect_acransac_batch(np.zeros((0, 2)), np.zeros((0, 3)), [0], scs[0]["K"]) == []


def test_batch_resection_recovers_the_poses(gpu):
    """400 views at once: every clean problem is localised to its true camera."""
    specs = [(300 + 7 * (i % 40), 0.3 + 0.01 * (i % 30)) for i in range(400)]
    scs, off, x2d, X3d = _batch_problems(specs, seed0=2000)
    got = gpu.resect_acransac_batch(x2d, X3d, off, scs[0]["K"], 4096, seed=3)
    for p, sc in enumerate(scs):
        assert got[p]["found"], p
        M = np.linalg.inv(sc["K"]) @ got[p]["P"]
        M /= np.cbrt(np.linalg.det(M[:, :3]))
        C = -M[:, :3].T @ M[:, 3]
        assert np.linalg.norm(C - (-sc["R"].T @ sc["t"])) < 0.05, p


@pytest.mark.parametrize("sequential", [False, True])
def test_acransac_inliers_agree_with_opencv_anchor(gpu, sequential):
    """Independent anchor (tests/golden/make_golden_anchors.py): the final inlier sets of the device
    AC-RANSAC against cv2.solvePnPRansac(P3P)'s at the threshold the fp64 restatement estimated, and
    against the planted truth.  OpenCV keeps an unrefined minimal model of its own sampling, so its
    set is a little smaller and nearly contained in ours."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    a = dict(np.load(os.path.join(root, "tests", "golden", "anchors_golden.npz")))
    for k, (N, outl, seed) in enumerate(a["resect_cases"]):
        sc = synth.resection_scene(int(N), int(seed), outlier_frac=float(outl))
        r = gpu.resect_acransac(sc["x2d"], sc["X3d"], sc["K"], max_iter=4096, seed=1, sequential=sequential)
        assert r["found"]
        mine = np.zeros(int(N), bool); mine[r["inliers"]] = True
        cv = a["resect_%d_cv2_inliers" % k]
        # the threshold is estimated, not given: it must land near the restatement's
        assert abs(r["error_max"] - float(a["resect_%d_threshold_px" % k])) < 0.5
        assert (cv & mine).sum() / cv.sum() >= 0.95
        assert (cv & mine).sum() / (cv | mine).sum() >= 0.88
        assert (sc["inlier_mask"] & mine).sum() / (sc["inlier_mask"] | mine).sum() >= 0.95
