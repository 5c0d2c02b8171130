"""GPU parity suite for guided matching (hulo_guided_match: bGuided_matching of hulo::geometricMatch)
against the CPU oracle.  The epipolar gate is decided by the same fp64 expression tree on both
sides and the rest is integer work, so the match lists must be IDENTICAL."""
import math

import numpy as np
import pytest

from sfmlocalization_b200 import synth

pytestmark = pytest.mark.gpu


def scene(V=5, feats=900, seed=41):
    sc = synth.localization_scene(V, feats, 1200, 10, seed, track_frac=0.7)
    off = sc["seg_offsets"]
    segs = [sc["rows"][int(off[k]):int(off[k + 1])] for k in range(V)]
    xys = [sc["map_xy"][int(off[k]):int(off[k + 1])] for k in range(V)]
    return sc, segs, xys


def models(orc, segs, xys, pairs, seed=3):
    Fs, thr = [], []
    for (I, J) in pairs:
        oi, oj = orc.match_pair(segs[I], segs[J], 0.7)
        r = orc.fmatrix_acransac(xys[I][oi], xys[J][oj], synth.IMAGE_WH, synth.IMAGE_WH, 4.0, 200, seed + I * 100 + J)
        assert r["ok"]
        Fs.append(r["F"]); thr.append(r["error_max"] ** 2)
    return np.array(Fs), np.array(thr)


def test_identical_to_oracle_over_a_batch_of_pairs(gpu, orc):
    sc, segs, xys = scene()
    pairs = [(0, 1), (0, 3), (2, 4), (3, 1), (1, 2), (4, 0)]
    Fs, thr = models(orc, segs, xys, pairs)
    db = gpu.db(sc["rows"], sc["seg_offsets"])
    try:
        off, gi, gj = gpu.guided_match(db, sc["map_xy"], pairs, Fs, thr)
        total = 0
        for p, (I, J) in enumerate(pairs):
            wi, wj = orc.guided_match(Fs[p], xys[I], segs[I], xys[J], segs[J], thr[p], 0.36)
            a, b = int(off[p]), int(off[p + 1])
            assert gi[a:b].tolist() == wi.tolist() and gj[a:b].tolist() == wj.tolist(), (I, J)
            total += len(wi)
        assert total > 300 and int(off[-1]) == total == len(gi)
        # capacity protocol: too small a buffer reports the size needed, nothing else changes
        import ctypes as C
        from sfmlocalization_b200 import _lib
        n = C.c_size_t(0)
        oi = np.zeros(4, np.uint32); oj = np.zeros(4, np.uint32); o = np.zeros(len(pairs) + 1, np.uint64)
        P = np.ascontiguousarray(pairs, np.uint32)
        st = gpu.lib.hulo_guided_match(gpu.h, db.h, sc["map_xy"].ctypes.data_as(C.c_void_p), P.ctypes.data_as(C.c_void_p),
                                       len(pairs), np.ascontiguousarray(Fs).ctypes.data_as(C.c_void_p),
                                       thr.ctypes.data_as(C.c_void_p), 0.36, 1, o.ctypes.data_as(C.c_void_p),
                                       oi.ctypes.data_as(C.c_void_p), oj.ctypes.data_as(C.c_void_p), 4, C.byref(n))
        assert st == _lib.ERR_CAPACITY and n.value == total
    finally:
        db.free()


@pytest.mark.parametrize("thr_px,ratio", [(0.5, 0.36), (2.0, 0.8), (25.0, 0.36), (math.inf, 0.5)])
def test_gate_width_and_ratio(gpu, orc, thr_px, ratio):
    """From a needle-thin gate to none at all (every j is a candidate: plain 2-NN + ratio)."""
    sc, segs, xys = scene(V=3, feats=700, seed=43)
    pairs = [(0, 1), (2, 0)]
    Fs, _ = models(orc, segs, xys, pairs)
    thr = np.full(2, thr_px ** 2)
    db = gpu.db(sc["rows"], sc["seg_offsets"])
    try:
        off, gi, gj = gpu.guided_match(db, sc["map_xy"], pairs, Fs, thr, dist_ratio=ratio)
        for p, (I, J) in enumerate(pairs):
            wi, wj = orc.guided_match(Fs[p], xys[I], segs[I], xys[J], segs[J], thr[p], ratio)
            a, b = int(off[p]), int(off[p + 1])
            assert gi[a:b].tolist() == wi.tolist() and gj[a:b].tolist() == wj.tolist()
    finally:
        db.free()


def test_duplicate_positions_empty_images_and_degenerate_models(gpu, orc):
    sc, segs, xys = scene(V=3, feats=500, seed=47)
    # image 0 gets 40 features twice (same position, same descriptor): the second of each is a
    # duplicate 4-tuple and is dropped; image 3 is empty; image 4 has a single feature
    rows = np.concatenate([segs[0], segs[0][:40], segs[1], segs[2], synth.random_rows(1, 5)])
    xy = np.concatenate([xys[0], xys[0][:40], xys[1], xys[2], np.array([[10.0, 20.0]])])
    off = np.array([0, 540, 1040, 1540, 1540, 1541], np.uint64)
    segs2 = [rows[int(off[k]):int(off[k + 1])] for k in range(5)]
    xys2 = [xy[int(off[k]):int(off[k + 1])] for k in range(5)]
    Fs, thr = models(orc, segs, xys, [(0, 1)])
    pairs = [(0, 1), (3, 1), (1, 3), (4, 1), (1, 4), (0, 1)]
    F6 = np.tile(Fs[0], (6, 1, 1)); t6 = np.tile(thr[0], 6)
    F6[5] = 0.0                                    # a zero model: every line is degenerate, nothing matches
    db = gpu.db(rows, off)
    try:
        o, gi, gj = gpu.guided_match(db, xy, pairs, F6, t6)
        o2, gi2, gj2 = gpu.guided_match(db, xy, pairs, F6, t6, dedup=False)
        for p, (I, J) in enumerate(pairs):
            wi, wj = orc.guided_match(F6[p], xys2[I], segs2[I], xys2[J], segs2[J], t6[p], 0.36)
            a, b = int(o[p]), int(o[p + 1])
            assert gi[a:b].tolist() == wi.tolist() and gj[a:b].tolist() == wj.tolist(), p
            ui, uj = orc.guided_match(F6[p], xys2[I], segs2[I], xys2[J], segs2[J], t6[p], 0.36, dedup=False)
            a, b = int(o2[p]), int(o2[p + 1])
            assert gi2[a:b].tolist() == ui.tolist() and gj2[a:b].tolist() == uj.tolist(), p
        assert int(o2[1]) > int(o[1]) > 50                     # duplicates existed and were dropped
        assert int(o[1]) == int(o[5]) == int(o[6])             # empty / single-feature / zero-model pairs: nothing
    finally:
        db.free()


def test_adversarial_geometry(gpu, orc):
    """Geometry that stresses the gate: lines along the axes, diagonal, through a corner, with the
    epipole far outside the image, a degenerate model; features off the image on every side and
    non-finite positions; gates from sub-pixel to 'everything'; images of one feature, of 40
    features within 20 px, one spanning 60000 px, 300 features at one point: the match lists stay
    those of the CPU restatement."""
    rng = np.random.default_rng(11)
    n_img = 6
    sizes = [700, 900, 1, 40, 800, 600]
    rows = [synth.random_rows(n, 500 + k) for k, n in enumerate(sizes)]
    # descriptors of every image near those of image 0's, so that nearest / second nearest are meaningful
    base = rows[0]
    for k in range(1, n_img):
        src = base[rng.integers(0, len(base), sizes[k])].copy()
        flip = rng.random(src.shape) < 0.02
        rows[k] = src ^ (flip * rng.integers(1, 255, src.shape)).astype(np.uint8)
        rows[k][:, 61:] = 0
    xys = [np.stack([rng.uniform(-300, 2300, n), rng.uniform(-200, 1300, n)], axis=1) for n in sizes]
    xys[1][::50] = np.nan
    xys[1][7] = [np.inf, 3.0]
    xys[3] = rng.uniform(0, 20, (sizes[3], 2))                 # everything in one cell
    xys[4][:, 0] *= 30.0                                       # 60000 px wide: the cell size grows
    xys[5][:300] = xys[5][0]                                   # 300 features at one point
    off = np.zeros(n_img + 1, np.uint64); off[1:] = np.cumsum(sizes)
    allrows = np.concatenate(rows); allxy = np.concatenate(xys)

    def cross(e):
        return np.array([[0, -e[2], e[1]], [e[2], 0, -e[0]], [-e[1], e[0], 0]], float)
    Fs = [
        cross([1.0, 0.0, 0.0]),                    # lines through (inf, 0): horizontal lines y = const
        cross([0.0, 1.0, 0.0]),                    # vertical lines
        cross([960.0, 540.0, 1.0]),                # pencil through the image centre: every slope
        cross([-5000.0, 300.0, 1.0]),              # epipole far to the left
        cross([0.0, 0.0, 1.0]),                    # lines through the corner (0, 0)
        np.array([[0, 0, 0], [0, 0, 0], [0, 0, 1.0]]),    # l = (0, 0, 1): degenerate, matches nothing
        np.array([[0, 0, 1e-9], [0, 0, 1.0], [0, 0, -640.0]]),   # one fixed, almost horizontal line
        rng.normal(size=(3, 3)),
    ]
    pairs, Fl, thr = [], [], []
    for (I, J) in [(0, 1), (1, 0), (0, 4), (4, 0), (0, 5), (5, 1), (0, 3), (3, 0), (0, 2), (2, 0), (1, 5)]:
        for F in Fs:
            for t in (0.25, 16.0, 2500.0, 1e12):
                pairs.append((I, J)); Fl.append(F); thr.append(t)
    Fl = np.array(Fl); thr = np.array(thr)
    db = gpu.db(allrows, off)
    try:
        o, gi, gj = gpu.guided_match(db, allxy, pairs, Fl, thr, dist_ratio=0.8)
    finally:
        db.free()
    total = 0
    for p, (I, J) in enumerate(pairs):
        wi, wj = orc.guided_match(Fl[p], xys[I], rows[I], xys[J], rows[J], thr[p], 0.8)
        a, b = int(o[p]), int(o[p + 1])
        assert gi[a:b].tolist() == wi.tolist() and gj[a:b].tolist() == wj.tolist(), (p, I, J, thr[p])
        total += len(wi)
    assert total > 2000

