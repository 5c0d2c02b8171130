"""CPU suite for the C++ host layer above the C-ABI (csrc/host): file formats, pair lists,
pair-list sharding, track propagation.  None of this needs a GPU."""
import json
import os

import numpy as np

from sfmlocalization_b200 import synth
from tests import hostlib


def test_desc_file_round_trip(tmp_path):
    rows61 = np.ascontiguousarray(synth.random_rows(37, 1)[:, :61])
    p = str(tmp_path / "a.desc")
    assert hostlib.write_desc(p, rows61) == 0                 # saveAKAZEBin pads 61 -> 64
    raw = open(p, "rb").read()
    assert len(raw) == 8 + 37 * 64 and int(np.frombuffer(raw[:8], np.uint64)[0]) == 37
    back = hostlib.read_desc(p)
    assert back.shape == (37, 64) and np.array_equal(back[:, :61], rows61) and (back[:, 61:] == 0).all()
    q = str(tmp_path / "b.desc")
    hostlib.write_desc_numpy(q, rows61)                       # independent writer, same bytes
    assert open(q, "rb").read() == raw
    hostlib.write_desc_numpy(q, np.zeros((0, 64), np.uint8))
    assert hostlib.read_desc(q).shape == (0, 64)
    assert hostlib.read_desc(str(tmp_path / "missing.desc")) is None


def test_pair_generators():
    ids = [0, 1, 2, 5, 9]
    want = [(a, b) for k, a in enumerate(ids) for b in ids[k + 1:]]
    assert hostlib.all_pairs(ids).tolist() == [list(p) for p in want]
    want_v = [(ids[a], ids[b]) for a in range(5) for b in range(a + 1, min(5, a + 3))]
    assert hostlib.video_pairs(ids, 2).tolist() == [list(p) for p in want_v]
    assert len(hostlib.video_pairs(ids, 0)) == 0


def ref_remove_dup(pairs):
    """python restatement of SfMDataUtils.cpp:168-190 (in-place normalisation included)."""
    pairs = [list(p) for p in pairs]
    dup = []
    for i in range(len(pairs) - 1, 0, -1):
        for j in range(i - 1, -1, -1):
            if pairs[j][0] > pairs[j][1]:
                pairs[j] = [pairs[j][1], pairs[j][0]]
            if pairs[i] == pairs[j] or pairs[i] == [pairs[j][1], pairs[j][0]]:
                dup.append(i)
                break
    for d in dup:
        del pairs[d]
    return pairs


def test_remove_dup_pairs_matches_reference_algorithm():
    rng = np.random.default_rng(4)
    for _ in range(20):
        p = rng.integers(0, 6, size=(rng.integers(1, 25), 2))
        assert hostlib.remove_dup_pairs(p).tolist() == ref_remove_dup(p.tolist())


def test_partition_pairs_covers_and_balances():
    rng = np.random.default_rng(5)
    rows = rng.integers(500, 6000, size=40)
    pairs = [(a, b) for a in range(40) for b in range(a + 1, 40)]
    cost = np.array([rows[a] * rows[b] for a, b in pairs], np.float64)
    for world in (2, 3, 8):
        parts = [hostlib.partition_pairs(pairs, rows, r, world) for r in range(world)]
        allpos = np.sort(np.concatenate(parts))
        assert np.array_equal(allpos, np.arange(len(pairs)))             # every pair exactly once
        loads = np.array([cost[p].sum() for p in parts])
        assert loads.max() / loads.mean() < 1.02                         # balanced by n_I * n_J
        assert all((np.diff(p) > 0).all() for p in parts)
    assert np.array_equal(hostlib.partition_pairs(pairs, rows, 0, 1), np.arange(len(pairs)))


def test_track_propagation_matches_oracle(orc):
    rng = np.random.default_rng(6)
    V, n = 7, 40
    feat = [n] * (V - 1)
    m_off, m_i, m_j = [0], [], []
    for f in range(V - 1):
        src = np.sort(rng.choice(n, size=25, replace=False)); dst = rng.permutation(n)[:25]
        m_i += src.tolist(); m_j += dst.tolist(); m_off.append(len(m_i))
    for max_dist in (2, 3, 5, 20):
        got = hostlib.propagate_tracks(V, max_dist, feat, m_off, m_i, m_j)
        f, t, i, j = orc.track_propagate(V, max_dist, feat, m_off, m_i, m_j)
        want = sorted(zip(f.tolist(), t.tolist(), i.tolist(), j.tolist()), key=lambda r: (r[0], r[1]))
        # the C++ layer returns map order (pair ascending), inside a pair emission order
        assert [tuple(r) for r in got.tolist()] == want


def test_match_file_round_trip(tmp_path):
    src = tmp_path / "in.txt"
    src.write_text("0 1\n2\n3 4\n5 6\n0 2\n0\n7 9\n1\n10 11\n")
    dst = tmp_path / "out.txt"
    assert hostlib.lib().hulo_host_matches_roundtrip(str(src).encode(), str(dst).encode()) == 3
    assert hostlib.parse_matches(str(dst)) == {(0, 1): [(3, 4), (5, 6)], (0, 2): [], (7, 9): [(10, 11)]}
    assert dst.read_text() == src.read_text()


def test_views_from_sfm_data_json(tmp_path):
    views = [{"key": k, "value": {"polymorphic_id": 1073741824, "ptr_wrapper": {"id": 2147483649 + k, "data": {
        "local_path": "/", "filename": "img%03d.jpg" % (3 * k), "width": 1920, "height": 1080, "id_view": k,
        "id_intrinsic": 0, "id_pose": k}}}} for k in range(5)]
    doc = {"sfm_data_version": "0.2", "root_path": "/data", "views": views,
           "intrinsics": [{"key": 0, "value": {"filename": "not-a-view", "id_view": 99}}], "extrinsics": []}
    p = tmp_path / "sfm_data.json"
    p.write_text(json.dumps(doc, indent=4))
    import ctypes as C
    ids = np.zeros(16, np.uint64); names = np.zeros((16, 64), np.uint8)
    n = hostlib.lib().hulo_host_views_from_sfm_data(str(p).encode(), ids.ctypes.data_as(C.c_void_p),
                                                    names.ctypes.data_as(C.c_void_p), C.c_ulonglong(16),
                                                    C.c_ulonglong(64))
    assert n == 5 and ids[:5].tolist() == [0, 1, 2, 3, 4]
    assert bytes(names[2]).split(b"\0")[0] == b"img006.jpg"


def test_sfm_data_json_reader(tmp_path):
    """loadSfMData: the cereal layout of OpenMVG 1.x (views / intrinsics / extrinsics / structure)."""
    sc = synth.localization_scene(5, 60, 200, 10, 3)
    names = ["frame%04d" % k for k in range(5)]
    p = str(tmp_path / "sfm_data.json")
    lms = hostlib.write_sfm_data(p, sc, names, disto=[0.05, -0.01, 0.002])
    r = hostlib.load_sfm_data(p)
    assert r["counts"].tolist() == [5, 1, 5, len(lms), len(sc["obs_view"])]
    assert np.allclose(r["first_X"], sc["landmark_X"][lms[0]])
    assert np.allclose(r["intrinsic0"], [sc["K"][0, 0], sc["K"][0, 2], sc["K"][1, 2], 1920, 1080, 0.05, -0.01, 0.002])
    assert (r["view_wh"] == [1920, 1080]).all()
    assert hostlib.load_sfm_data(str(tmp_path / "missing.json")) is None
    (tmp_path / "bad.json").write_text("{\"views\": [")
    assert hostlib.load_sfm_data(str(tmp_path / "bad.json")) is None


def test_undistortion_inverts_the_radial_model():
    """Intrinsic::get_ud_pixel (Pinhole_Intrinsic_Radial_K3): distort(undistort(p)) == p."""
    f, cx, cy, k = 1800.0, 960.0, 540.0, (0.08, -0.02, 0.004)
    for x, y in [(100.0, 80.0), (960.0, 540.0), (1900.0, 1000.0), (300.5, 777.25)]:
        ux, uy = hostlib.undistort(f, cx, cy, k, x, y)
        a, b = (ux - cx) / f, (uy - cy) / f
        r2 = a * a + b * b
        c = 1 + k[0] * r2 + k[1] * r2 ** 2 + k[2] * r2 ** 3
        assert abs(f * a * c + cx - x) < 1e-4 and abs(f * b * c + cy - y) < 1e-4
    assert np.array_equal(hostlib.undistort(f, cx, cy, (0, 0, 0), 12.5, 99.0), [12.5, 99.0])   # plain pinhole


def test_opencv_yaml_matrix_and_feat_reader(tmp_path):
    import cv2
    A = np.arange(12, dtype=np.float64).reshape(3, 4) * 0.37 - 1.5
    fs = cv2.FileStorage(str(tmp_path / "A.yml"), cv2.FILE_STORAGE_WRITE)
    fs.write("other", np.eye(2)); fs.write("A", A); fs.release()
    got = hostlib.read_cv_matrix(str(tmp_path / "A.yml"), "A")
    assert got.shape == (3, 4) and np.allclose(got, A)
    assert hostlib.read_cv_matrix(str(tmp_path / "A.yml"), "missing") is None
    xy = np.array([[1.5, 2.25], [1000.125, 33.0]])
    hostlib.write_feat(str(tmp_path / "a.feat"), xy)
    assert np.array_equal(hostlib.read_feat(str(tmp_path / "a.feat")), xy)
    assert hostlib.read_feat(str(tmp_path / "none.feat")) is None


def test_sfm_data_rewrite_keeps_everything_but_the_poses(tmp_path):
    """saveSfMDataPoses (what hulo_ba_resect writes, adjust_sfm_data.cpp:152-155): the extrinsics
    are replaced, every other member comes back with every number spelled as it was read."""
    import json
    sc = synth.localization_scene(5, 40, 60, 30, 23)
    names = ["v%02d" % k for k in range(5)]
    p = str(tmp_path / "sfm_data.json")
    hostlib.write_sfm_data(p, sc, names, disto=[0.05, -0.01, 0.002])
    doc = json.load(open(p))
    doc["root_path"] = "/d\u00e9p\u00f4t \U0001F600/\"q\"\\x\ttab"   # escapes: 2- and 4-byte UTF-8, a surrogate pair, quotes
    json.dump(doc, open(p, "w"))                                   # ensure_ascii: written as \uXXXX escapes
    sums = hostlib.sfm_observation_sums(p)
    off = sc["seg_offsets"].astype(np.int64)
    want = sc["map_xy"][off[sc["obs_view"]] + sc["obs_feat"]].sum(axis=0)
    assert np.allclose(sums, want, rtol=1e-12)                     # Observation::x is read
    rng = np.random.default_rng(5)
    ids = np.array([0, 2, 3, 9], np.uint64)                         # a pose may be new (9) or dropped (1, 4)
    R = rng.normal(size=(4, 3, 3)); Cc = rng.normal(size=(4, 3)) * 1e3
    out = str(tmp_path / "out.json")
    assert hostlib.save_sfm_poses(p, out, ids, R, Cc)
    a, b = json.load(open(p)), json.load(open(out))
    assert list(a.keys()) == list(b.keys())
    for k in a:
        if k != "extrinsics":
            assert a[k] == b[k], k
    assert [e["key"] for e in b["extrinsics"]] == [0, 2, 3, 9]
    for e, r, c in zip(b["extrinsics"], R, Cc):
        assert np.array_equal(np.array(e["value"]["rotation"]), r)  # 17 significant digits: exact round trip
        assert np.array_equal(np.array(e["value"]["center"]), c)
    # numbers that were not touched keep their spelling
    txt = open(out).read()
    assert repr(float(sc["landmark_X"][0][0])) in txt
    # and the file loads again
    r = hostlib.load_sfm_data(out)
    assert r["counts"][2] == 4 and r["counts"][3] == json.load(open(p))["structure"].__len__()
    assert not hostlib.save_sfm_poses(str(tmp_path / "missing.json"), out, ids, R, Cc)
