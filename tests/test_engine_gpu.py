"""GPU suite for the end-to-end query localisation (hulo_engine_localize): the deterministic
stages (matching, view filter, 2D-3D assembly) bit-exact against the oracle composition, the
resection statistically (AC-RANSAC draws differ) against ground truth."""
import numpy as np
import pytest

from sfmlocalization_b200 import synth
from sfmlocalization_b200.gpu import LocalizeEngine

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["int", "tc"])
def engine(request, gpu):
    """Every test of this module runs with the 2-NN arithmetic on the integer pipes (K1) and on the
    tensor cores (K1t, item mode): same bit-exact expectations."""
    gpu.set_knn_engine(request.param)
    yield request.param
    gpu.set_knn_engine("auto")


def oracle_assembly(orc, sc, views, ratio, min_putative):
    off = sc["seg_offsets"]
    m_view, m_i, m_j, m_d = [], [], [], []
    for v in sorted(set(views)):
        a = sc["rows"][int(off[v]):int(off[v + 1])]
        oi, oj, od = orc.match_view_to_query(a, sc["q_desc"], ratio)
        if len(oi) < min_putative:
            continue
        m_view += [v] * len(oi); m_i += oi.tolist(); m_j += oj.tolist(); m_d += od.tolist()
    order = np.lexsort((sc["obs_feat"], sc["obs_view"]))
    return orc.match_set(m_view, m_i, m_j, m_view, m_j, m_d, sc["obs_view"][order], sc["obs_feat"][order],
                         sc["obs_landmark"][order].astype(np.int64), len(sc["q_desc"]))


@pytest.mark.parametrize("seed,views", [(1, None), (2, None), (3, [9, 2, 3, 4, 11, 17])])
def test_localize_matches_oracle_pipeline(gpu, orc, seed, views):
    sc = synth.localization_scene(24, 800, 4000, 900, seed)
    eng = LocalizeEngine(gpu, sc["rows"], sc["seg_offsets"], sc["obs_view"], sc["obs_feat"], sc["obs_landmark"],
                         sc["landmark_X"], sc["K"], ratio=0.6)
    try:
        r = eng.localize(sc["q_desc"], sc["q_xy"], views=views, seed=5)
    finally:
        eng.close()
    vs = list(range(24)) if views is None else views
    wj, wl = oracle_assembly(orc, sc, vs, 0.6, 16)
    assert r["corr_qfeat"].tolist() == wj.tolist()
    assert r["corr_landmark"].tolist() == wl.tolist()
    # the assembled pairs are overwhelmingly the true ones
    truth = sc["q_truth"][r["corr_qfeat"]]
    assert (truth == r["corr_landmark"]).mean() > 0.9
    assert r["localized"]
    assert np.linalg.norm(r["center"] - sc["center"]) < 0.05
    assert np.abs(r["R"] - sc["R"]).max() < 5e-3
    assert len(r["inliers"]) > 10
    good = truth[r["inliers"]] == r["corr_landmark"][r["inliers"]]
    assert good.mean() > 0.95


def test_localize_fails_cleanly_without_overlap(gpu):
    sc = synth.localization_scene(6, 300, 1000, 200, 7, query_inlier_frac=0.0)
    eng = LocalizeEngine(gpu, sc["rows"], sc["seg_offsets"], sc["obs_view"], sc["obs_feat"], sc["obs_landmark"],
                         sc["landmark_X"], sc["K"])
    try:
        r = eng.localize(sc["q_desc"], sc["q_xy"])
        assert not r["localized"] and len(r["corr_qfeat"]) <= 8
        r = eng.localize(np.zeros((0, 64), np.uint8), np.zeros((0, 2)))
        assert not r["localized"]
    finally:
        eng.close()


def test_localize_respects_thresholds(gpu):
    sc = synth.localization_scene(10, 400, 1500, 300, 9)
    kw = dict(rows=sc["rows"], seg_offsets=sc["seg_offsets"], obs_view=sc["obs_view"], obs_feat=sc["obs_feat"],
              obs_landmark=sc["obs_landmark"], landmark_X=sc["landmark_X"], K=sc["K"])
    eng = LocalizeEngine(gpu, **kw, min_putative=10 ** 6)      # every view dropped (LocalizeEngine.cc:428-434)
    r = eng.localize(sc["q_desc"], sc["q_xy"]); eng.close()
    assert not r["localized"] and len(r["corr_qfeat"]) == 0
    eng = LocalizeEngine(gpu, **kw, min_inliers=10 ** 6)       # resection runs, result rejected (:560)
    r = eng.localize(sc["q_desc"], sc["q_xy"]); eng.close()
    assert not r["localized"] and len(r["inliers"]) > 10


def test_match_to_queries_equals_single_calls(gpu):
    """The batched matching pass returns, per query image, exactly what hulo_match_to_query does."""
    sc = synth.localization_scene(12, 600, 2500, 500, 21)
    qs = [sc["q_desc"]] + [synth.extra_query(sc, n, 30 + k)["q_desc"] for k, n in enumerate((300, 1, 700))]
    qs.insert(2, np.zeros((0, 64), np.uint8))                      # an image without descriptors
    off = np.zeros(len(qs) + 1, np.uint64); off[1:] = np.cumsum([q.shape[0] for q in qs])
    db = gpu.db(sc["rows"], sc["seg_offsets"])
    try:
        for views in (None, [7, 2, 3, 4]):
            import os
            for budget in (None, "9000"):                          # second pass: several sub-batches
                if budget:
                    os.environ["HULO_QUERY_BATCH_ROWS"] = budget
                b = gpu.match_to_queries(db, np.concatenate(qs), off, 0.6, views=views, cap=64)
                os.environ.pop("HULO_QUERY_BATCH_ROWS", None)
                k = 0
                for q, rows in enumerate(qs):
                    s = gpu.match_to_query(db, rows, 0.6, views=None if views is None else np.array(views, np.uint32))
                    n = len(s["i"])
                    assert (b["query"][k:k + n] == q).all()
                    assert np.array_equal(b["view"][k:k + n], s["view"]) and np.array_equal(b["i"][k:k + n], s["i"])
                    assert np.array_equal(b["j"][k:k + n], s["j"]) and np.array_equal(b["d0"][k:k + n], s["d0"])
                    assert np.array_equal(b["counts"][q], s["view_counts"])
                    k += n
                assert k == len(b["i"]) and k > 200
    finally:
        db.free()


def test_localize_batch(gpu):
    sc = synth.localization_scene(16, 700, 3000, 800, 23)
    extra = [synth.extra_query(sc, 800, 40 + k) for k in range(5)]
    eng = LocalizeEngine(gpu, sc["rows"], sc["seg_offsets"], sc["obs_view"], sc["obs_feat"], sc["obs_landmark"],
                         sc["landmark_X"], sc["K"], ratio=0.6)
    try:
        descs = [sc["q_desc"]] + [e["q_desc"] for e in extra]
        xys = [sc["q_xy"]] + [e["q_xy"] for e in extra]
        centres = [sc["center"]] + [e["center"] for e in extra]
        b = eng.localize_batch(descs, xys, seed=3)
        for q in range(len(descs)):
            s = eng.localize(descs[q], xys[q], seed=3 + q)        # the batch uses seed + q
            assert b["localized"][q] and s["localized"]
            assert b["n_corr"][q] == len(s["corr_qfeat"]) and b["n_inliers"][q] == len(s["inliers"])
            assert np.allclose(b["center"][q], s["center"], atol=1e-12) and np.allclose(b["R"][q], s["R"], atol=1e-12)
            assert np.linalg.norm(b["center"][q] - centres[q]) < 0.05
    finally:
        eng.close()


def oracle_geometric_assembly(orc, sc, views, ratio, min_putative, rounds, precision, seed):
    """Putative matching -> view filter -> F-matrix filter per (view, query) pair -> assembly over
    the geometric matches with featDist from the putative ones (LocalizeEngine.cc:423-497)."""
    off = sc["seg_offsets"]
    w, h = synth.IMAGE_WH
    g_view, g_i, g_j, f_view, f_j, f_d = [], [], [], [], [], []
    p = 0
    kept = []
    for v in sorted(set(views)):
        a = sc["rows"][int(off[v]):int(off[v + 1])]
        oi, oj, od = orc.match_view_to_query(a, sc["q_desc"], ratio)
        if len(oi) < min_putative:
            continue
        r = orc.fmatrix_acransac(sc["map_xy"][int(off[v]) + oi], sc["q_xy"][oj], (w, h), (w, h), precision, rounds,
                                 seed + 77 + 1000003 * p)
        p += 1
        if not r["ok"]:
            continue
        kept.append(v)
        inl = r["inliers"]
        g_view += [v] * len(inl); g_i += oi[inl].tolist(); g_j += oj[inl].tolist()
        f_view += [v] * len(oi); f_j += oj.tolist(); f_d += od.tolist()
    order = np.lexsort((sc["obs_feat"], sc["obs_view"]))
    wj, wl = orc.match_set(g_view, g_i, g_j, f_view, f_j, f_d, sc["obs_view"][order], sc["obs_feat"][order],
                           sc["obs_landmark"][order].astype(np.int64), len(sc["q_desc"]))
    return wj, wl, kept


@pytest.mark.parametrize("seed,rounds", [(1, 25), (2, 200)])
def test_localize_with_geometric_filter(gpu, orc, seed, rounds):
    sc = synth.localization_scene(24, 800, 4000, 900, seed)
    eng = LocalizeEngine(gpu, sc["rows"], sc["seg_offsets"], sc["obs_view"], sc["obs_feat"], sc["obs_landmark"],
                         sc["landmark_X"], sc["K"], ratio=0.6)
    try:
        plain = eng.localize(sc["q_desc"], sc["q_xy"], seed=5)
        with pytest.raises(Exception):
            eng.configure_geometric(True, rounds, 4.0)            # keypoints not set yet: loud failure
        eng.set_keypoints(sc["map_xy"], sc["view_wh"], synth.IMAGE_WH)
        eng.configure_geometric(True, rounds, 4.0)
        r = eng.localize(sc["q_desc"], sc["q_xy"], seed=5)
        eng.configure_geometric(False)
        again = eng.localize(sc["q_desc"], sc["q_xy"], seed=5)
    finally:
        eng.close()
    assert again["corr_qfeat"].tolist() == plain["corr_qfeat"].tolist()
    wj, wl, kept = oracle_geometric_assembly(orc, sc, range(24), 0.6, 16, rounds, 4.0, 5)
    assert len(kept) >= 12
    got = set(zip(r["corr_qfeat"].tolist(), r["corr_landmark"].tolist()))
    want = set(zip(wj.tolist(), wl.tolist()))
    assert len(got & want) >= 0.97 * len(got | want)
    # the filter removes wrong pairs: what is left is cleaner than the putative assembly
    truth = sc["q_truth"][r["corr_qfeat"]]
    truth_plain = sc["q_truth"][plain["corr_qfeat"]]
    assert (truth == r["corr_landmark"]).mean() >= (truth_plain == plain["corr_landmark"]).mean()
    assert (truth == r["corr_landmark"]).mean() > 0.97
    assert r["localized"] and np.linalg.norm(r["center"] - sc["center"]) < 0.05
    assert r["times_ms"][3] > 0 and plain["times_ms"][3] == 0


def composed_guided_assembly(gpu, orc, sc, seed, rounds, precision, ratio=0.6, min_putative=16):
    """The engine's guided path rebuilt from the library's stand-alone entry points: putative
    matching, F-matrix filter per (view, query) pair with the engine's seeds, guided matching of
    the surviving pairs on a table that holds the query as one more image, then the assembly of
    SfMDataUtils.cpp:59-125 (restated on the CPU) over the guided matches, featDist from the
    putative ones."""
    off = sc["seg_offsets"].astype(np.int64)
    V = len(off) - 1
    w, h = synth.IMAGE_WH
    mdb = gpu.db(sc["rows"], sc["seg_offsets"])
    m = gpu.match_to_query(mdb, sc["q_desc"], ratio)
    mdb.free()
    sel = [v for v in range(V) if m["view_counts"][v] >= min_putative]
    xI, xJ, poff = [], [], [0]
    for v in sel:
        k = m["view"] == v
        xI.append(sc["map_xy"][off[v] + m["i"][k]]); xJ.append(sc["q_xy"][m["j"][k]])
        poff.append(poff[-1] + int(k.sum()))
    f = gpu.geometric_filter(np.concatenate(xI), np.concatenate(xJ), poff, [[w, h, w, h]] * len(sel), precision, rounds,
                             seed + 77)
    valid = [p for p in range(len(sel)) if f["valid"][p]]
    both = gpu.db(np.concatenate([sc["rows"], sc["q_desc"]]), np.append(sc["seg_offsets"], len(sc["rows"]) + len(sc["q_desc"])))
    goff, gi, gj = gpu.guided_match(both, np.concatenate([sc["map_xy"], sc["q_xy"]]), [(sel[p], V) for p in valid],
                                    f["F"][valid], f["error_max"][valid] ** 2, 0.36)
    both.free()
    g_view, f_view, f_j, f_d = [], [], [], []
    for n, p in enumerate(valid):
        g_view += [sel[p]] * int(goff[n + 1] - goff[n])
        k = m["view"] == sel[p]
        f_view += [sel[p]] * int(k.sum()); f_j += m["j"][k].tolist(); f_d += m["d0"][k].tolist()
    order = np.lexsort((sc["obs_feat"], sc["obs_view"]))
    wj, wl = orc.match_set(g_view, gi.tolist(), gj.tolist(), f_view, f_j, f_d, sc["obs_view"][order],
                           sc["obs_feat"][order], sc["obs_landmark"][order].astype(np.int64), len(sc["q_desc"]))
    return wj, wl, len(valid), len(gi)


@pytest.mark.parametrize("seed", [1, 4])
def test_localize_with_guided_matching(gpu, orc, seed):
    """mGuidedMatching (LocalizeEngine.cc:458 -> MatchUtils.cpp:407-416) inside the per-query engine:
    exactly the correspondences the stand-alone entry points give when chained by hand."""
    sc = synth.localization_scene(24, 800, 4000, 900, seed)
    eng = LocalizeEngine(gpu, sc["rows"], sc["seg_offsets"], sc["obs_view"], sc["obs_feat"], sc["obs_landmark"],
                         sc["landmark_X"], sc["K"], ratio=0.6)
    try:
        with pytest.raises(Exception):
            eng.set_guided_matching(True)                        # needs the geometric filter
        eng.set_keypoints(sc["map_xy"], sc["view_wh"], synth.IMAGE_WH)
        eng.configure_geometric(True, 25, 4.0)
        unguided = eng.localize(sc["q_desc"], sc["q_xy"], seed=5)
        eng.set_guided_matching(True)
        r = eng.localize(sc["q_desc"], sc["q_xy"], seed=5)
        r2 = eng.localize(sc["q_desc"], sc["q_xy"], seed=5)      # cached position groups: same answer
        b = eng.localize_batch([sc["q_desc"], sc["q_desc"][:400], sc["q_desc"]], [sc["q_xy"], sc["q_xy"][:400], sc["q_xy"]],
                               seed=5)
        eng.configure_geometric(False)                           # takes guided matching with it
        plain = eng.localize(sc["q_desc"], sc["q_xy"], seed=5)
    finally:
        eng.close()
    wj, wl, n_valid, n_guided = composed_guided_assembly(gpu, orc, sc, 5, 25, 4.0)
    assert n_valid >= 12 and n_guided > 200
    assert r["corr_qfeat"].tolist() == wj.tolist() and r["corr_landmark"].tolist() == wl.tolist()
    assert r2["corr_qfeat"].tolist() == wj.tolist()
    assert int(b["n_corr"][0]) == len(wj)                        # batched entry point: query rows at an offset
    assert b["localized"][0] and np.allclose(b["center"][0], r["center"])
    assert r["localized"] and np.linalg.norm(r["center"] - sc["center"]) < 0.05
    assert plain["times_ms"][3] == 0
    # guided matching only re-pairs query features the putative matching already reached
    assert set(r["corr_qfeat"].tolist()) <= set(plain["corr_qfeat"].tolist()) | set(unguided["corr_qfeat"].tolist())
    truth = sc["q_truth"][r["corr_qfeat"]]
    assert (truth == r["corr_landmark"]).mean() > 0.97


def test_engine_sequential_schedule_is_the_letter_exact_driver(gpu, orc):
    """hulo_engine_set_resection_schedule(SEQUENTIAL): the engine's pose and inlier list are those of
    hulo_resect_acransac_sequential (the reference's AC-RANSAC loop kept to the letter, trace-checked
    against the oracle in test_resect_gpu.py) on the correspondences it assembled."""
    sc = synth.localization_scene(24, 800, 4000, 900, 4)
    eng = LocalizeEngine(gpu, sc["rows"], sc["seg_offsets"], sc["obs_view"], sc["obs_feat"], sc["obs_landmark"],
                         sc["landmark_X"], sc["K"], ratio=0.6)
    try:
        eng.set_resection_schedule("sequential")
        r = eng.localize(sc["q_desc"], sc["q_xy"], seed=9)
        eng.set_resection_schedule("batched")
        rb = eng.localize(sc["q_desc"], sc["q_xy"], seed=9)
    finally:
        eng.close()
    x2d = sc["q_xy"][r["corr_qfeat"]]; X3d = sc["landmark_X"][r["corr_landmark"]]
    want = gpu.resect_acransac(x2d, X3d, sc["K"], max_iter=4096, seed=9, sequential=True)
    assert r["localized"] and want["found"]
    assert r["inliers"].tolist() == want["inliers"].tolist()
    _, R, c = gpu.pose_from_projection(want["P"])
    assert np.abs(r["R"] - R).max() < 1e-12 and np.abs(r["center"] - c).max() < 1e-12
    # and the default schedule lands on the same pose statistically
    assert rb["localized"] and np.linalg.norm(rb["center"] - r["center"]) < 0.02
    o = orc.acransac(x2d, X3d, sc["K"], max_iter=4096, seed=9)
    assert o["ok"] and abs(len(o["inliers"]) - len(r["inliers"])) <= max(5, len(o["inliers"]) // 20)
