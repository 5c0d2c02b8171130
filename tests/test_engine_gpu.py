"""GPU suite for the end-to-end query localisation (hulo_engine_localize): the deterministic
stages (matching, view filter, 2D-3D assembly) bit-exact against the oracle composition, the
resection statistically (AC-RANSAC draws differ) against ground truth."""
import numpy as np
import pytest

from sfmlocalization_b200 import synth
from sfmlocalization_b200.gpu import LocalizeEngine

pytestmark = pytest.mark.gpu


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
