"""GPU suite for the file-level surface (what the reference's Python scripts drive): the
hulo_ext_match CLI reads .desc files and writes matches.putative.txt, and the C++ hulo::
entry points behind it; results equal the oracle composition on the same files."""
import json
import os
import subprocess

import numpy as np
import pytest

from sfmlocalization_b200 import synth
from tests import hostlib

pytestmark = pytest.mark.gpu


def make_matchdir(tmp_path, n_images, rows_per_image, seed, jitter=60, with_json=True):
    rows, off = synth.image_collection(n_images, rows_per_image, seed, overlap=0.6, jitter=jitter)
    d = tmp_path / "matches"
    d.mkdir()
    names = ["frame%04d" % k for k in range(n_images)]
    segs = []
    for k in range(n_images):
        seg = rows[int(off[k]):int(off[k + 1])]
        hostlib.write_desc_numpy(str(d / (names[k] + ".desc")), np.ascontiguousarray(seg[:, :61]))
        segs.append(seg)
    if with_json:
        views = [{"key": k, "value": {"ptr_wrapper": {"data": {"local_path": "/", "filename": names[k] + ".jpg",
                                                               "width": 1920, "height": 1080, "id_view": k}}}}
                 for k in range(n_images)]
        (d / "sfm_data.json").write_text(json.dumps({"root_path": "/x", "views": views, "intrinsics": []}))
    return d, segs


def run_cli(*args):
    r = subprocess.run([hostlib.CLI] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def oracle_pairs(orc, segs, pairs, ratio):
    out = {}
    for I, J in pairs:
        oi, oj = orc.match_pair(segs[I], segs[J], ratio)
        if len(oi):
            out.setdefault((I, J), [])
            out[(I, J)] += list(zip(oi.tolist(), oj.tolist()))
    return out


def test_all_pairs_mode(tmp_path, orc):
    d, segs = make_matchdir(tmp_path, 7, 500, 1)
    run_cli(d, "-f=0.7", "-r=500", "-mm=30", "-gm")                 # unknown reference flags tolerated
    got = hostlib.parse_matches(str(d / "matches.putative.txt"))
    pairs = [(a, b) for a in range(7) for b in range(a + 1, 7)]
    assert got == oracle_pairs(orc, segs, pairs, 0.7)
    assert sum(len(v) for v in got.values()) > 300


def test_video_window_and_pair_file_modes(tmp_path, orc):
    d, segs = make_matchdir(tmp_path, 6, 400, 2)
    run_cli(d, "-f=0.6", "-v=2", "--out=" + str(d / "v.txt"))
    pairs = [(a, b) for a in range(6) for b in range(a + 1, min(6, a + 3))]
    assert hostlib.parse_matches(str(d / "v.txt")) == oracle_pairs(orc, segs, pairs, 0.6)
    pf = d / "pairs.txt"
    pf.write_text("0 3\n4 1\n2 5\n")
    run_cli(d, "-f=0.8", "-p=" + str(pf), "--out=" + str(d / "p.txt"))
    assert hostlib.parse_matches(str(d / "p.txt")) == oracle_pairs(orc, segs, [(0, 3), (4, 1), (2, 5)], 0.8)


def test_tracking_mode(tmp_path, orc):
    V = 6
    d, segs = make_matchdir(tmp_path, V, 450, 3, jitter=0)
    run_cli(d, "-f=0.7", "-mf=4")
    got = hostlib.parse_matches(str(d / "matches.putative.txt"))
    want = oracle_pairs(orc, segs, [(f, f + 1) for f in range(V - 1)], 0.7)
    m_off, m_i, m_j = [0], [], []
    for f in range(V - 1):
        for (i, j) in want.get((f, f + 1), []):
            m_i.append(i); m_j.append(j)
        m_off.append(len(m_i))
    f, t, i, j = orc.track_propagate(V, 4, [len(s) for s in segs[:-1]], m_off, m_i, m_j)
    for a, b, c, e in zip(f.tolist(), t.tolist(), i.tolist(), j.tolist()):
        want.setdefault((a, b), []).append((c, e))
    got = {k: v for k, v in got.items() if v}       # propagation creates empty (f, f+1) keys like the reference
    assert got == want
    assert any(k[1] - k[0] >= 2 for k in got)


def test_pair_list_sharded_over_two_ranks(tmp_path, orc):
    """Reconstruction matching shards the pair list with no collective: the two ranks' files
    together equal the single-process file."""
    d, segs = make_matchdir(tmp_path, 8, 300, 4, jitter=120, with_json=False)
    (d / "views.txt").write_text("".join("%d frame%04d.jpg\n" % (k, k) for k in range(8)))
    run_cli(d, "-f=0.7", "--views=" + str(d / "views.txt"), "--out=" + str(d / "one.txt"))
    for r in range(2):
        run_cli(d, "-f=0.7", "--views=" + str(d / "views.txt"), "--out=" + str(d / "two.txt"), "--rank=%d" % r,
                "--world=2", "--device=0")
    one = hostlib.parse_matches(str(d / "one.txt"))
    a = hostlib.parse_matches(str(d / "two.txt.rank0")); b = hostlib.parse_matches(str(d / "two.txt.rank1"))
    assert not (set(a) & set(b))
    merged = dict(a); merged.update(b)
    assert merged == one
    assert len(a) > 5 and len(b) > 5


GEO_SEED = 0x5EED      # hulo::g_geometricSeed


def test_geometric_stage_writes_matches_f(tmp_path, orc):
    """ExtFeatAndMatch's second stage (computeFeaturesAndMatches.cpp:194-246): pairs below -mm are
    dropped, the rest go through the F-matrix filter; matches.f.txt holds the inliers in ACRANSAC's
    order.  Same composition on the oracle, same per-pair sampler seed."""
    V = 6
    sc = synth.localization_scene(V, 700, 900, 10, 31, track_frac=0.7)
    off = sc["seg_offsets"]
    d = tmp_path / "matches"
    d.mkdir()
    segs, xys = [], []
    for k in range(V):
        seg = sc["rows"][int(off[k]):int(off[k + 1])]
        xy = sc["map_xy"][int(off[k]):int(off[k + 1])]
        hostlib.write_desc_numpy(str(d / ("frame%04d.desc" % k)), np.ascontiguousarray(seg[:, :61]))
        with open(d / ("frame%04d.feat" % k), "w") as f:
            for x, y in xy:
                f.write("%r %r 1.0 0.0\n" % (float(x), float(y)))
        segs.append(seg); xys.append(xy)
    views = [{"key": k, "value": {"ptr_wrapper": {"data": {"local_path": "/", "filename": "frame%04d.jpg" % k,
                                                           "width": 1920, "height": 1080, "id_view": k}}}}
             for k in range(V)]
    (d / "sfm_data.json").write_text(json.dumps({"root_path": "/x", "views": views, "intrinsics": []}))
    out = run_cli(d, "-f=0.7", "-r=200", "-mm=40", "-g=4.0")
    assert "geometric matching skipped" not in out
    put = hostlib.parse_matches(str(d / "matches.putative.txt"))
    geo = hostlib.parse_matches(str(d / "matches.f.txt"))
    pairs = [(a, b) for a in range(V) for b in range(a + 1, V)]
    want_put = oracle_pairs(orc, segs, pairs, 0.7)
    assert put == want_put
    exact = 0
    n_valid = 0
    for (I, J), m in want_put.items():
        if len(m) < 40:
            assert (I, J) not in geo
            continue
        m = np.array(m)
        r = orc.fmatrix_acransac(xys[I][m[:, 0]], xys[J][m[:, 1]], (1920, 1080), (1920, 1080), 4.0, 200,
                                 GEO_SEED + 1000003 * (I * 1000003 + J))
        assert ((I, J) in geo) == r["ok"]
        if not r["ok"]:
            continue
        n_valid += 1
        want = [tuple(x) for x in m[r["inliers"]].tolist()]
        got = geo[(I, J)]
        assert len(set(got) & set(want)) >= 0.8 * len(set(got) | set(want))
        exact += int(set(got) == set(want))
    assert n_valid >= 5 and exact >= n_valid - 2


def test_geometric_stage_skipped_without_feat_files(tmp_path):
    d, segs = make_matchdir(tmp_path, 4, 300, 9)
    out = run_cli(d, "-f=0.7", "-mm=5")
    assert "geometric matching skipped" in out and not (d / "matches.f.txt").exists()
    out = run_cli(d, "-f=0.7", "--putative-only")
    assert "Geometric" not in out


def test_guided_matching_flag(tmp_path, orc):
    """-gm: the inliers of every surviving pair are replaced by guided matches over all features of
    the two images (ReconstructParam.py:70-71 switches it on for reconstruction)."""
    V = 4
    sc = synth.localization_scene(V, 700, 900, 10, 33, track_frac=0.7)
    off = sc["seg_offsets"]
    d = tmp_path / "matches"
    d.mkdir()
    segs, xys = [], []
    for k in range(V):
        seg = sc["rows"][int(off[k]):int(off[k + 1])]
        xy = sc["map_xy"][int(off[k]):int(off[k + 1])]
        hostlib.write_desc_numpy(str(d / ("frame%04d.desc" % k)), np.ascontiguousarray(seg[:, :61]))
        hostlib.write_feat(str(d / ("frame%04d.feat" % k)), xy)
        segs.append(seg); xys.append(xy)
    views = [{"key": k, "value": {"ptr_wrapper": {"data": {"local_path": "/", "filename": "frame%04d.jpg" % k,
                                                           "width": 1920, "height": 1080, "id_view": k}}}}
             for k in range(V)]
    (d / "sfm_data.json").write_text(json.dumps({"root_path": "/x", "views": views, "intrinsics": []}))
    run_cli(d, "-f=0.7", "-r=200", "-mm=40", "-g=4.0", "-gm")
    geo = hostlib.parse_matches(str(d / "matches.f.txt"))
    put = hostlib.parse_matches(str(d / "matches.putative.txt"))
    assert len(geo) == 6
    for (I, J), got in geo.items():
        m = np.array(put[(I, J)])
        r = orc.fmatrix_acransac(xys[I][m[:, 0]], xys[J][m[:, 1]], (1920, 1080), (1920, 1080), 4.0, 200,
                                 GEO_SEED + 1000003 * (I * 1000003 + J))
        assert r["ok"]
        wi, wj = orc.guided_match(r["F"], xys[I], segs[I], xys[J], segs[J], r["error_max"] ** 2, 0.36)
        want = set(zip(wi.tolist(), wj.tolist()))
        assert len(set(got) & want) >= 0.97 * len(set(got) | want) and len(got) > 100
        assert [g[0] for g in got] == sorted(g[0] for g in got)            # ascending i
