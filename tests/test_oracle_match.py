"""CPU suite: the oracle's matching restatement against the committed golden vectors
(OpenCV exact matchers, tests/golden/make_golden.py) and against independent numpy
restatements of the reference's post filters."""
import numpy as np
import pytest

from sfmlocalization_b200 import synth

INT_MAX = 2**31 - 1


def np_knn2(A, B):
    """numpy lexsort restatement: (distance, index) ascending."""
    A64 = np.zeros((A.shape[0], 64), np.uint8); A64[:, :A.shape[1]] = A
    B64 = np.zeros((B.shape[0], 64), np.uint8); B64[:, :B.shape[1]] = B
    x = np.bitwise_xor(A64[:, None, :], B64[None, :, :])
    d = np.unpackbits(x, axis=2).sum(axis=2).astype(np.int32)
    idx = np.full((A.shape[0], 2), -1, np.int32)
    dist = np.full((A.shape[0], 2), INT_MAX, np.int32)
    for i in range(A.shape[0]):
        order = np.lexsort((np.arange(B.shape[0]), d[i]))[:2]
        idx[i, :len(order)] = order
        dist[i, :len(order)] = d[i, order]
    return idx, dist


@pytest.mark.parametrize("name", ["tie", "pl", "w61", "dup"])
def test_knn2_matches_opencv_golden(orc, golden, name):
    idx, dist = orc.knn2(golden[name + "_A"], golden[name + "_B"])
    assert np.array_equal(idx, golden[name + "_idx"])
    assert np.array_equal(dist, golden[name + "_dist"])


def test_knn2_planted_targets_found(orc, golden):
    idx, _ = orc.knn2(golden["pl_A"], golden["pl_B"])
    hit = golden["pl_target"] >= 0
    assert hit.sum() > 50
    assert np.array_equal(idx[hit, 0], golden["pl_target"][hit])


def test_knn2_61_equals_64_padded(orc, golden):
    A64 = orc.pad_rows(golden["w61_A"]); B64 = orc.pad_rows(golden["w61_B"])
    assert A64.shape == (64, 64) and (A64[:, 61:] == 0).all()
    i61, d61 = orc.knn2(golden["w61_A"], golden["w61_B"])
    i64, d64 = orc.knn2(A64, B64)
    assert np.array_equal(i61, i64) and np.array_equal(d61, d64)


@pytest.mark.parametrize("nB", [0, 1, 2, 3])
def test_knn2_degenerate_database(orc, nB):
    A = synth.random_rows(5, 1); B = synth.random_rows(nB, 2)
    idx, dist = orc.knn2(A, B)
    ri, rd = np_knn2(A, B)
    assert np.array_equal(idx, ri) and np.array_equal(dist, rd)
    if nB == 0:
        assert (idx == -1).all() and (dist == INT_MAX).all()
    if nB == 1:
        assert (idx[:, 1] == -1).all() and (dist[:, 1] == INT_MAX).all()


def test_knn2_random_vs_numpy(orc):
    A, B, _ = synth.descriptor_sets(40, 300, 5)
    idx, dist = orc.knn2(A, B)
    ri, rd = np_knn2(A, B)
    assert np.array_equal(idx, ri) and np.array_equal(dist, rd)


def test_lsh_reference_convention(golden):
    """The reference's approximate matcher marks a missing neighbour with -1 / INT_MAX
    (guarded at MatchUtils.cpp:115, 203, 349); the exact matcher never does when nB >= 2."""
    li, ld = golden["lsh_idx"], golden["lsh_dist"]
    miss = li == -1
    assert (ld[miss] == INT_MAX).all()
    assert (golden["pl_idx"] >= 0).all()
    # LSH finds most planted first neighbours but is not the parity target
    hit = golden["pl_target"] >= 0
    assert (li[hit, 0] == golden["pl_target"][hit]).mean() > 0.8


def test_ratio_cases(orc, golden):
    for d0, d1, ratio, expect in golden["ratio_cases"]:
        assert orc.ratio_pass(int(d0), int(d1), float(ratio)) == bool(expect), (d0, d1, ratio)


def ref_pair_filter(idx2, dist2, ratio):
    """Line-by-line python restatement of MatchUtils.cpp:111-150 (O(n^2) loop kept)."""
    n = idx2.shape[0]
    NONE = -1
    m = [0] * n
    for i in range(n):
        with np.errstate(divide="ignore", invalid="ignore"):
            q = (np.float32(0.0) + np.float32(dist2[i, 0])) / np.float32(dist2[i, 1])
        if q < np.float32(ratio):
            if dist2[i, 1] < INT_MAX:
                m[i] = int(idx2[i, 0])
        else:
            m[i] = NONE
    for i in range(n - 1):
        if m[i] == NONE:
            continue
        dup = False
        for j in range(i + 1, n):
            if m[i] == m[j]:
                m[j] = NONE
                dup = True
        if dup:
            m[i] = NONE
    return [(i, m[i]) for i in range(n - 1) if m[i] != NONE]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_pair_filter_vs_python(orc, seed):
    rows, off = synth.image_collection(2, 180, seed, overlap=0.6)
    A, B = rows[:int(off[1])], rows[int(off[1]):]
    # force duplicates: several rows of A are copies of one another
    A = A.copy(); A[5] = A[3]; A[7] = A[3]; A[-1] = A[20]
    idx2, dist2 = orc.knn2(A, B)
    oi, oj = orc.pair_filter(idx2, dist2, 0.7)
    ref = ref_pair_filter(idx2, dist2, 0.7)
    assert list(zip(oi.tolist(), oj.tolist())) == ref
    assert len(ref) > 10
    assert (oi < A.shape[0] - 1).all()          # last row never emitted (MatchUtils.cpp:146)
    assert len(set(oj.tolist())) == len(oj)     # one-to-one
    mi, mj = orc.match_pair(A, B, 0.7)
    assert np.array_equal(mi, oi) and np.array_equal(mj, oj)


def test_pair_filter_unfound_second_neighbour_quirk(orc):
    """d1 == INT_MAX with a passing float ratio keeps the value-initialised index 0
    (MatchUtils.cpp:111-117)."""
    idx2 = np.array([[4, -1], [2, 3], [7, -1]], np.int32)
    dist2 = np.array([[10, INT_MAX], [10, 100], [5, INT_MAX]], np.int32)
    oi, oj = orc.pair_filter(idx2, dist2, 0.6)
    ref = ref_pair_filter(idx2, dist2, 0.6)
    assert list(zip(oi.tolist(), oj.tolist())) == ref == [(1, 2)]   # rows 0 and 2 collide on index 0


def test_match_pair_skips_tiny_images(orc):
    A = synth.random_rows(1, 1); B = synth.random_rows(50, 2)
    assert len(orc.match_pair(A, B, 0.8)[0]) == 0
    assert len(orc.match_pair(B, A, 0.8)[0]) == 0


def test_match_view_to_query(orc):
    Bq = synth.random_rows(150, 9)
    A, target = synth.plant_matches(synth.random_rows(400, 10), Bq, 11, frac=0.4)
    i, j, d0 = orc.match_view_to_query(A, Bq, 0.6)
    idx2, dist2 = orc.knn2(A, Bq)
    keep = [k for k in range(400) if orc.ratio_pass(dist2[k, 0], dist2[k, 1], 0.6)]
    assert i.tolist() == keep
    assert np.array_equal(j, idx2[keep, 0]) and np.array_equal(d0, dist2[keep, 0])
    hit = np.nonzero(target >= 0)[0]
    assert set(hit.tolist()) <= set(keep)
    # a single query row can never pass: d1 == INT_MAX (MatchUtils.cpp:349)
    assert len(orc.match_view_to_query(A, Bq[:1], 0.6)[0]) == 0


def ref_track(n_frames, max_dist, feat_number, matches):
    """python restatement of MatchUtils.cpp:239-276."""
    tp = []
    for f in range(n_frames - 1):
        t = [-1] * feat_number[f]
        for (i, j) in matches.get((f, f + 1), []):
            t[i] = j
        tp.append(t)
    out = []
    for f in range(n_frames - 1):
        for to in range(f + 2, min(f + max_dist, n_frames)):
            for i in range(len(tp[f])):
                t = tp[f][i]
                if t != -1:
                    nx = tp[to - 1][t]
                    tp[f][i] = nx
                    if nx != -1:
                        out.append((f, to, i, nx))
    return out


def test_track_propagate(orc):
    rng = np.random.default_rng(3)
    V, n = 6, 30
    feat_number = [n] * (V - 1)
    matches = {}
    m_off = [0]; m_i = []; m_j = []
    for f in range(V - 1):
        src = np.sort(rng.choice(n, size=18, replace=False))
        dst = rng.permutation(n)[:18]
        matches[(f, f + 1)] = list(zip(src.tolist(), dst.tolist()))
        m_i += src.tolist(); m_j += dst.tolist(); m_off.append(len(m_i))
    for max_dist in (2, 3, 4, 10):
        f, t, i, j = orc.track_propagate(V, max_dist, feat_number, m_off, m_i, m_j)
        assert list(zip(f.tolist(), t.tolist(), i.tolist(), j.tolist())) == ref_track(V, max_dist, feat_number, matches)


def test_match_set_closest_descriptor_wins(orc):
    # two views both match query feature 3; the smaller featDist wins, ties keep the first view
    m_view = [0, 0, 1, 1, 2]; m_i = [5, 6, 2, 9, 1]; m_j = [3, 4, 3, 8, 3]
    fd_view = [0, 0, 1, 1, 2]; fd_j = [3, 4, 3, 8, 3]; fd_d = [40, 30, 25, 50, 25]
    lm_view = [0, 0, 1, 2]; lm_feat = [5, 6, 2, 1]; lm_id = [100, 101, 200, 300]   # (1,9) has no landmark
    j, lm = orc.match_set(m_view, m_i, m_j, fd_view, fd_j, fd_d, lm_view, lm_feat, lm_id, 10)
    assert j.tolist() == [3, 4] and lm.tolist() == [200, 101]
