"""GPU suite for hulo_ba_resect / hulo::resectViews: the resection stage of the reference's
OpenMVG_BA tool (adjust_sfm_data.cpp:91-155) -- every view re-resected against the structure it
observes, all views in one hulo_resect_acransac_batch call, poses written back to sfm_data.json."""
import json
import subprocess

import numpy as np
import pytest

from sfmlocalization_b200 import synth
from tests import hostlib

pytestmark = pytest.mark.gpu
K_EQ = np.array([[1865.0, 0.0, 1043.21], [0.0, 1865.0, 644.65], [0.0, 0.0, 1.0]])


def run(*args):
    return subprocess.run([hostlib.CLI_BA_RESECT] + [str(a) for a in args], capture_output=True, text=True, timeout=600)


def make_scene(tmp_path, V=24, seed=41):
    sc = synth.localization_scene(V, 500, 2500, 10, seed, K=K_EQ)
    names = ["frame%04d" % k for k in range(V)]
    p = tmp_path / "sfm_data.json"
    hostlib.write_sfm_data(str(p), sc, names)
    # forget the poses: the tool has to find them from the 2D-3D pairs alone
    doc = json.load(open(p))
    for e in doc["extrinsics"]:
        e["value"] = {"rotation": np.eye(3).tolist(), "center": [0.0, 0.0, 0.0]}
    json.dump(doc, open(p, "w"))
    return sc, p


def test_views_are_resected_to_their_true_poses(tmp_path, gpu):
    sc, p = make_scene(tmp_path)
    out = tmp_path / "sfm_data_b4bd.json"
    r = run(p, out, "-r=0", "--seed=7")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Resected 24 of 24 views" in r.stdout
    doc, src = json.load(open(out)), json.load(open(p))
    for k in src:
        if k != "extrinsics":
            assert doc[k] == src[k], k
    assert [e["key"] for e in doc["extrinsics"]] == list(range(24))
    # the same 2D-3D pairs in ascending landmark id, the same per-view seeds, through the Python
    # binding of the same entry point: the poses in the file are those bits
    off_rows = sc["seg_offsets"].astype(np.int64)
    x2d, X3d, offs = [], [], [0]
    for v in range(24):
        m = sc["obs_view"] == v
        order = np.argsort(sc["obs_landmark"][m], kind="stable")
        x2d.append(sc["map_xy"][off_rows[v] + sc["obs_feat"][m]][order])
        X3d.append(sc["landmark_X"][sc["obs_landmark"][m]][order])
        offs.append(offs[-1] + int(m.sum()))
    seeds = np.uint64(7) + np.uint64(1000003) * np.arange(24, dtype=np.uint64)
    want = gpu.resect_acransac_batch(np.concatenate(x2d), np.concatenate(X3d), offs, K_EQ, 4096, seeds=seeds)
    for v, e in enumerate(doc["extrinsics"]):
        R = np.array(e["value"]["rotation"]); C = np.array(e["value"]["center"])
        C_true = -sc["view_R"][v].T @ sc["view_t"][v]
        assert np.linalg.norm(C - C_true) < 0.05, v
        assert np.abs(R - sc["view_R"][v]).max() < 5e-3, v
        assert want[v]["found"]
        Kd, Rw, Cw = gpu.pose_from_projection(want[v]["P"])
        assert np.array_equal(R, Rw) and np.array_equal(C, Cw), v
        assert np.allclose(Kd, K_EQ, rtol=2e-3, atol=3.0)


def test_views_with_too_few_points_keep_their_pose(tmp_path):
    sc, p = make_scene(tmp_path, V=6, seed=43)
    doc = json.load(open(p))
    # view 2 keeps 10 observations (not MORE than 10: skipped with the warning), view 4 gets pure noise
    kept = 0
    rng = np.random.default_rng(1)
    for lm in doc["structure"]:
        obs = []
        for o in lm["value"]["observations"]:
            if o["key"] == 2:
                kept += 1
                if kept > 10:
                    continue
            if o["key"] == 4:
                o["value"]["x"] = [float(rng.uniform(0, 1920)), float(rng.uniform(0, 1080))]
            obs.append(o)
        lm["value"]["observations"] = obs
    doc["extrinsics"][2]["value"]["center"] = [1.0, 2.0, 3.0]
    doc["extrinsics"][4]["value"]["center"] = [4.0, 5.0, 6.0]
    json.dump(doc, open(p, "w"))
    out = tmp_path / "out.json"
    r = run(p, out)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Warning: there is/are frames with too few matches." in r.stdout      # adjust_sfm_data.cpp:148-150
    assert "Resected 4 of 5 views (6 in the file)" in r.stdout
    res = json.load(open(out))
    assert res["extrinsics"][2]["value"]["center"] == [1.0, 2.0, 3.0]
    assert res["extrinsics"][4]["value"]["center"] == [4.0, 5.0, 6.0]
    for v in (0, 1, 3, 5):
        C_true = -sc["view_R"][v].T @ sc["view_t"][v]
        assert np.linalg.norm(np.array(res["extrinsics"][v]["value"]["center"]) - C_true) < 0.05


def test_usage_and_refusals(tmp_path):
    assert run().returncode == 1
    sc, p = make_scene(tmp_path, V=3, seed=44)
    r = run(p, tmp_path / "o.json", "-c=rst,rsti")
    assert r.returncode == 1 and "bundle adjustment" in r.stderr
    assert run(tmp_path / "missing.json", tmp_path / "o.json").returncode != 0
