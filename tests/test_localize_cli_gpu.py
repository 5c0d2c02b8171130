"""GPU suite for hulo_localize (the OpenMVGLocalization_AKAZE CLI from the putative matching on)
and the C++ hulo::LocalizeEngine behind it: sfm_data.json + .desc/.feat files in, per-query
result JSON out (keys filename, sfm_data, matches_dir, K, R, t, pair -- localization.cpp:100-144)."""
import json
import os
import subprocess

import numpy as np
import pytest

from sfmlocalization_b200 import synth
from sfmlocalization_b200.gpu import LocalizeEngine
from tests import hostlib

pytestmark = pytest.mark.gpu
K_EQ = np.array([[1865.0, 0.0, 1043.21], [0.0, 1865.0, 644.65], [0.0, 0.0, 1.0]])   # OpenMVG pinhole: one focal


def make_project(tmp_path, seed=5, V=14, n_queries=3):
    sc = synth.localization_scene(V, 700, 3000, 800, seed, K=K_EQ)
    sfm = tmp_path / "sfm"; mdir = tmp_path / "matches"; qdir = tmp_path / "query"; out = tmp_path / "out"
    for d in (sfm, mdir, qdir, out):
        d.mkdir()
    names = ["frame%04d" % k for k in range(V)]
    off = sc["seg_offsets"]
    for k in range(V):
        hostlib.write_desc_numpy(str(mdir / (names[k] + ".desc")), sc["rows"][int(off[k]):int(off[k + 1]), :61])
        hostlib.write_feat(str(mdir / (names[k] + ".feat")), sc["map_xy"][int(off[k]):int(off[k + 1])])
    lm_ids = hostlib.write_sfm_data(str(sfm / "sfm_data.json"), sc, names)
    queries = [dict(q_desc=sc["q_desc"], q_xy=sc["q_xy"], center=sc["center"], R=sc["R"])]
    queries += [synth.extra_query(sc, 800, 60 + k) for k in range(n_queries - 1)]
    for k, q in enumerate(queries):
        hostlib.write_desc_numpy(str(qdir / ("q%03d.desc" % k)), q["q_desc"][:, :61])
        hostlib.write_feat(str(qdir / ("q%03d.feat" % k)), q["q_xy"])
    return sc, sfm, mdir, qdir, out, queries, lm_ids


def run(*args):
    r = subprocess.run([hostlib.CLI_LOCALIZE] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    return r


def test_folder_of_queries(tmp_path):
    sc, sfm, mdir, qdir, out, queries, lm_ids = make_project(tmp_path)
    r = run(qdir, sfm, mdir, out, "-f=0.6", "-r=25", "-g=4.0", "-w", "-k=0")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "localized 3 of 3" in r.stdout
    for k, q in enumerate(queries):
        res = json.loads((out / ("q%03d.json" % k)).read_text())
        assert res["filename"].endswith("q%03d.desc" % k) and res["sfm_data"].endswith("sfm_data.json")
        assert res["matches_dir"] == str(mdir)
        assert np.allclose(np.array(res["K"]), K_EQ, rtol=1e-5)
        assert np.linalg.norm(np.array(res["t"]) - q["center"]) < 0.05            # t = camera centre -R^T t
        assert np.abs(np.array(res["R"]) - q["R"]).max() < 5e-3
        pairs = np.array(res["pair"])
        assert len(pairs) > 10 and pairs[:, 0].max() < 800 and np.isin(pairs[:, 1], lm_ids).all()
        if k == 0:                                                                # truth known for the main query
            assert (sc["q_truth"][pairs[:, 0]] == pairs[:, 1]).mean() > 0.95


def test_cli_equals_c_abi_engine(tmp_path, gpu):
    """The C++ engine over the files gives what the C-ABI engine gives on the same arrays and seed."""
    sc, sfm, mdir, qdir, out, queries, lm_ids = make_project(tmp_path, seed=8, n_queries=1)
    r = run(qdir / "q000.desc", sfm, mdir, out, "-f=0.6", "-r=25", "--seed=4")
    assert r.returncode == 0, r.stdout + r.stderr
    res = json.loads((out / "q000.json").read_text())
    eng = LocalizeEngine(gpu, sc["rows"], sc["seg_offsets"], sc["obs_view"], sc["obs_feat"], sc["obs_landmark"],
                         sc["landmark_X"], K_EQ, ratio=np.float32(0.6))
    try:
        eng.set_keypoints(sc["map_xy"], sc["view_wh"], synth.IMAGE_WH)
        eng.configure_geometric(True, 25, 4.0)
        e = eng.localize(sc["q_desc"], sc["q_xy"], seed=4 + 1)                    # the CLI uses seed + image number
    finally:
        eng.close()
    assert e["localized"]
    assert np.allclose(np.array(res["t"]), e["center"], rtol=2e-6, atol=1e-6)     # JSON keeps 6 significant digits
    assert np.allclose(np.array(res["R"]), e["R"], rtol=2e-6, atol=1e-6)
    want = [[int(e["corr_qfeat"][i]), int(e["corr_landmark"][i])] for i in e["inliers"]]
    assert res["pair"] == want


def test_failure_writes_the_short_json_and_bad_input_fails(tmp_path):
    sc, sfm, mdir, qdir, out, queries, lm_ids = make_project(tmp_path, seed=9, n_queries=1)
    hostlib.write_desc_numpy(str(qdir / "noise.desc"), synth.random_rows(300, 77)[:, :61])
    rng = np.random.default_rng(1)
    hostlib.write_feat(str(qdir / "noise.feat"), rng.uniform(0, 1000, (300, 2)))
    r = run(qdir / "noise.desc", sfm, mdir, out)
    assert r.returncode == 0 and "Fail to estimate camera matrix" in r.stdout
    res = json.loads((out / "noise.json").read_text())
    assert sorted(res) == ["filename", "matches_dir", "sfm_data"]                 # localization.cpp:84-99
    assert run(qdir, tmp_path / "nowhere", mdir, out).returncode != 0             # unreadable sfm_data.json
    assert run(qdir, sfm).returncode != 0                                         # usage


def test_restricting_views_by_location(tmp_path):
    """-x -y -z -d keep the views whose centre is near a location (hulo::getLocalViews)."""
    sc, sfm, mdir, qdir, out, queries, lm_ids = make_project(tmp_path, seed=10, n_queries=1)
    c = sc["center"]
    r = run(qdir, sfm, mdir, out, "-x=%r" % c[0], "-y=%r" % c[1], "-z=%r" % c[2], "-d=100.0")
    assert r.returncode == 0 and "localized 1 of 1" in r.stdout
    r = run(qdir, sfm, mdir, out, "-x=1000.0", "-y=1000.0", "-z=1000.0", "-d=0.5")   # nothing nearby
    assert r.returncode == 0 and "localized 0 of 1" in r.stdout


def test_queries_sharded_over_processes(tmp_path):
    """--rank/--world: every process localises its share of the folder; together they cover it."""
    sc, sfm, mdir, qdir, out, queries, lm_ids = make_project(tmp_path, seed=12, n_queries=4)
    for r in range(2):
        res = run(qdir, sfm, mdir, out, "-r=25", "--rank=%d" % r, "--world=2", "--device=0")
        assert res.returncode == 0 and "localized 2 of 4" in res.stdout, res.stdout + res.stderr
    for k, q in enumerate(queries):
        j = json.loads((out / ("q%03d.json" % k)).read_text())
        assert np.linalg.norm(np.array(j["t"]) - q["center"]) < 0.05


def test_bow_preselection(tmp_path):
    """-k: the views are narrowed to the knn nearest in bag-of-features space (views' .bow files and
    the query's .bow) before matching; localisation still succeeds on the narrowed set."""
    sc, sfm, mdir, qdir, out, queries, lm_ids = make_project(tmp_path, seed=13, n_queries=1)
    rng = np.random.default_rng(2)
    bof = rng.random((14, 96)).astype(np.float32)
    for v in range(14):
        hostlib.save_mat_bin(str(mdir / ("frame%04d.bow" % v)), bof[v].reshape(-1, 1), 5)
    qv = bof[[3, 7, 9]].mean(axis=0)
    hostlib.save_mat_bin(str(qdir / "q000.bow"), qv.reshape(-1, 1), 5)
    r = run(qdir, sfm, mdir, out, "-r=25", "-k=6")
    assert r.returncode == 0 and "number of selected local views by bow : 6" in r.stdout, r.stdout + r.stderr
    assert "localized 1 of 1" in r.stdout
    j = json.loads((out / "q000.json").read_text())
    assert np.linalg.norm(np.array(j["t"]) - queries[0]["center"]) < 0.05
    r = run(qdir, sfm, mdir, out, "-r=25", "-k=50")                 # more than there are views: no narrowing
    assert r.returncode == 0 and "selected local views by bow" not in r.stdout and "localized 1 of 1" in r.stdout


def test_global_coordinates_through_the_a_matrix(tmp_path):
    """--amat (the server's aMatFile, LocalizeEngine.cc:113-144): the model is moved to global
    coordinates first, so the pose comes out there: centre = A [c; 1], R = R_local A_R^T."""
    import cv2
    sc, sfm, mdir, qdir, out, queries, lm_ids = make_project(tmp_path, seed=14, n_queries=1)
    rng = np.random.default_rng(3)
    Q = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    Q *= np.sign(np.linalg.det(Q))
    A = np.hstack([Q, np.array([[5.0], [-2.0], [1.5]])])          # a rigid motion keeps the intrinsics valid
    fs = cv2.FileStorage(str(tmp_path / "A.yml"), cv2.FILE_STORAGE_WRITE)
    fs.write("A", A); fs.release()
    r = run(qdir, sfm, mdir, out, "-r=25", "--amat=" + str(tmp_path / "A.yml"))
    assert r.returncode == 0 and "localized 1 of 1" in r.stdout, r.stdout + r.stderr
    j = json.loads((out / "q000.json").read_text())
    want_c = A[:, :3] @ queries[0]["center"] + A[:, 3]
    assert np.linalg.norm(np.array(j["t"]) - want_c) < 0.05
    assert np.abs(np.array(j["R"]) - queries[0]["R"] @ Q.T).max() < 5e-3


def test_radial_distortion_is_removed_before_matching_geometry(tmp_path):
    """A pinhole_radial_k3 camera: the .feat positions of views and query are distorted pixels; the
    engine undistorts them (get_ud_pixel) for the geometric filter and the resection."""
    k3 = (0.12, -0.05, 0.01)
    sc = synth.localization_scene(12, 700, 3000, 800, 15, K=K_EQ)
    f, cx, cy = K_EQ[0, 0], K_EQ[0, 2], K_EQ[1, 2]

    def distort(xy):
        a = (xy[:, 0] - cx) / f; b = (xy[:, 1] - cy) / f
        r2 = a * a + b * b
        c = 1 + k3[0] * r2 + k3[1] * r2 ** 2 + k3[2] * r2 ** 3
        return np.stack([f * a * c + cx, f * b * c + cy], axis=1)

    sfm = tmp_path / "sfm"; mdir = tmp_path / "matches"; qdir = tmp_path / "query"; out = tmp_path / "out"
    for d in (sfm, mdir, qdir, out):
        d.mkdir()
    names = ["frame%04d" % k for k in range(12)]
    off = sc["seg_offsets"]
    for k in range(12):
        hostlib.write_desc_numpy(str(mdir / (names[k] + ".desc")), sc["rows"][int(off[k]):int(off[k + 1]), :61])
        hostlib.write_feat(str(mdir / (names[k] + ".feat")), distort(sc["map_xy"][int(off[k]):int(off[k + 1])]))
    hostlib.write_sfm_data(str(sfm / "sfm_data.json"), sc, names, disto=k3)
    hostlib.write_desc_numpy(str(qdir / "q000.desc"), sc["q_desc"][:, :61])
    hostlib.write_feat(str(qdir / "q000.feat"), distort(sc["q_xy"]))
    r = run(qdir, sfm, mdir, out, "-r=25")
    assert r.returncode == 0 and "localized 1 of 1" in r.stdout, r.stdout + r.stderr
    j = json.loads((out / "q000.json").read_text())
    assert np.linalg.norm(np.array(j["t"]) - sc["center"]) < 0.05       # with the distortion left in: decimetres off
    assert np.abs(np.array(j["R"]) - sc["R"]).max() < 5e-3


def test_guided_matching_flag(tmp_path):
    """-gm (localization.cpp:82): the F-matrix filter is followed by guided matching."""
    sc, sfm, mdir, qdir, out, queries, lm_ids = make_project(tmp_path, seed=8, n_queries=2)
    r = run(qdir, sfm, mdir, out, "-f=0.6", "-r=25", "-g=4.0", "-gm")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "localized 2 of 2" in r.stdout
    for k, q in enumerate(queries):
        res = json.load(open(out / ("q%03d.json" % k)))
        assert np.linalg.norm(np.array(res["t"]) - q["center"]) < 0.05
