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
