"""GPU parity suite for the two matching entry points behind hulo::matchAKAZEToQuery and
hulo::matchAKAZE / trackAKAZE, through the C-ABI, bit-exact against the CPU oracle."""
import numpy as np
import pytest

from sfmlocalization_b200 import _lib, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["int", "tc", "tc8"])
def engine(request, gpu):
    """Every test of this module runs with the 2-NN arithmetic on the integer pipes (K1) and on the
    tensor cores (K1t, item mode): same bit-exact expectations."""
    gpu.set_knn_engine(request.param)
    yield request.param
    gpu.set_knn_engine("auto")


def oracle_query(orc, rows, off, views, query, ratio):
    out = dict(view=[], i=[], j=[], d0=[], counts=[])
    for pos, v in enumerate(views):
        a = rows[int(off[v]):int(off[v + 1])]
        oi, oj, od = orc.match_view_to_query(a, query, ratio)
        out["view"] += [pos] * len(oi); out["i"] += oi.tolist(); out["j"] += oj.tolist(); out["d0"] += od.tolist()
        out["counts"].append(len(oi))
    return out


def check_query(gpu, orc, rows, off, views, query, ratio):
    db = gpu.db(rows, off)
    try:
        m = gpu.match_to_query(db, query, ratio, views=None if views is None else np.array(views, np.uint32))
    finally:
        db.free()
    vs = list(range(len(off) - 1)) if views is None else list(views)
    want = oracle_query(orc, rows, off, vs, query, ratio)
    assert m["view"].tolist() == want["view"]
    assert m["i"].tolist() == want["i"] and m["j"].tolist() == want["j"] and m["d0"].tolist() == want["d0"]
    assert m["view_counts"].tolist() == want["counts"]
    return len(want["i"])


@pytest.mark.parametrize("ratio", [0.6, 0.7, 0.8])
def test_match_to_query_all_views(gpu, orc, ratio):
    rows, off = synth.image_collection(12, 700, 5, overlap=0.0, jitter=150)
    query, _ = synth.plant_matches(synth.random_rows(900, 6), rows, 7, frac=0.5)
    assert check_query(gpu, orc, rows, off, None, query, ratio) > 100


def test_match_to_query_view_subset_and_order(gpu, orc):
    rows, off = synth.image_collection(20, 300, 8, overlap=0.0, jitter=100)
    query, _ = synth.plant_matches(synth.random_rows(500, 9), rows, 10, frac=0.6)
    assert check_query(gpu, orc, rows, off, [3, 4, 5, 11, 17, 18], query, 0.6) > 20
    assert check_query(gpu, orc, rows, off, [17, 2, 9, 3], query, 0.6) > 10      # caller's order kept


def test_match_to_query_large_views_and_chunked_query(gpu, orc):
    """Views larger than one searcher tile and a query large enough to be split into chunks."""
    rows, off = synth.image_collection(5, 5000, 12, overlap=0.0, jitter=800)
    query, _ = synth.plant_matches(synth.random_rows(6000, 13), rows, 14, frac=0.3)
    assert check_query(gpu, orc, rows, off, None, query, 0.7) > 500


def test_match_to_query_degenerate(gpu, orc):
    rows, off = synth.image_collection(4, 50, 15, overlap=0.0)
    # one query row: d1 == INT_MAX, nothing passes (MatchUtils.cpp:349)
    assert check_query(gpu, orc, rows, off, None, synth.random_rows(1, 1), 0.8) == 0
    # two query rows, identical to map rows: d0 == 0 passes any ratio
    q = rows[[3, 120]].copy()
    assert check_query(gpu, orc, rows, off, None, q, 0.6) >= 2
    # empty query / empty view in the middle
    db = gpu.db(rows, off)
    m = gpu.match_to_query(db, np.zeros((0, 64), np.uint8), 0.6)
    db.free()
    assert len(m["i"]) == 0
    off2 = np.array([0, 50, 50, 120, 200], np.uint64)
    q, _ = synth.plant_matches(synth.random_rows(60, 2), rows, 3, frac=0.7)
    assert check_query(gpu, orc, rows, off2, None, q, 0.7) > 5


def test_match_to_query_capacity_protocol(gpu):
    rows, off = synth.image_collection(3, 200, 16, overlap=0.0)
    q, _ = synth.plant_matches(synth.random_rows(300, 4), rows, 5, frac=0.9)
    db = gpu.db(rows, off)
    with pytest.raises(_lib.HuloError) as e:
        gpu.match_to_query(db, q, 0.8, cap=3)
    db.free()
    assert e.value.status == _lib.ERR_CAPACITY


def oracle_pairs(orc, rows, off, pairs, ratio, flags):
    offs = [0]; oi_all = []; oj_all = []
    for I, J in pairs:
        A = rows[int(off[I]):int(off[I + 1])]; B = rows[int(off[J]):int(off[J + 1])]
        if flags == _lib.PAIR_REFERENCE:
            oi, oj = orc.match_pair(A, B, ratio)
        else:
            if A.shape[0] < 2 or B.shape[0] < 2:
                oi = oj = np.zeros(0, np.int32)
            else:
                idx2, dist2 = orc.knn2(A, B)
                keep = [k for k in range(A.shape[0]) if orc.ratio_pass(dist2[k, 0], dist2[k, 1], ratio)]
                m = idx2[keep, 0]
                if flags & _lib.PAIR_ONE_TO_ONE:
                    cnt = np.bincount(m, minlength=B.shape[0])
                    sel = [k for k, t in zip(keep, m) if cnt[t] == 1]
                else:
                    sel = keep
                if flags & _lib.PAIR_DROP_LAST:
                    sel = [k for k in sel if k != A.shape[0] - 1]
                oi = np.array(sel, np.int32); oj = idx2[sel, 0] if len(sel) else np.zeros(0, np.int32)
        oi_all += oi.tolist(); oj_all += oj.tolist(); offs.append(len(oi_all))
    return offs, oi_all, oj_all


@pytest.mark.parametrize("flags", [_lib.PAIR_REFERENCE, 0, _lib.PAIR_ONE_TO_ONE, _lib.PAIR_DROP_LAST])
def test_match_pairs_vs_oracle(gpu, orc, flags):
    rows, off = synth.image_collection(8, 600, 21, overlap=0.5, jitter=120)
    # duplicate rows inside an image so that several rows claim the same train row
    rows = rows.copy()
    a0 = int(off[2]); rows[a0 + 7] = rows[a0 + 3]; rows[a0 + 9] = rows[a0 + 3]
    rows[int(off[3]) - 1] = rows[a0 + 20]
    pairs = [(i, j) for i in range(8) for j in range(i + 1, 8)] + [(5, 1), (3, 3)]
    db = gpu.db(rows, off)
    try:
        po, oi, oj = gpu.match_pairs(db, pairs, 0.7, flags=flags)
    finally:
        db.free()
    wo, wi, wj = oracle_pairs(orc, rows, off, pairs, 0.7, flags)
    assert po.tolist() == wo
    assert oi.tolist() == wi and oj.tolist() == wj
    assert len(wi) > 500


def test_match_pairs_tiny_and_empty_images(gpu, orc):
    rows = synth.random_rows(300, 30)
    off = np.array([0, 1, 1, 100, 102, 300], np.uint64)       # sizes 1, 0, 99, 2, 198
    rows[100:102] = rows[150:152]                               # image 3 duplicates two rows of image 4
    pairs = [(0, 2), (1, 2), (2, 0), (2, 1), (3, 4), (4, 3), (2, 4), (4, 2)]
    db = gpu.db(rows, off)
    try:
        po, oi, oj = gpu.match_pairs(db, pairs, 0.8)
    finally:
        db.free()
    wo, wi, wj = oracle_pairs(orc, rows, off, pairs, 0.8, _lib.PAIR_REFERENCE)
    assert po.tolist() == wo and oi.tolist() == wi and oj.tolist() == wj
    assert po[4] == 0            # pairs with an image of < 2 rows are skipped (MatchUtils.cpp:99-101)


def test_match_pairs_batches_and_capacity(gpu, orc, monkeypatch):
    rows, off = synth.image_collection(6, 400, 33, overlap=0.6)
    pairs = [(i, i + 1) for i in range(5)] + [(i, i + 2) for i in range(4)]
    wo, wi, wj = oracle_pairs(orc, rows, off, pairs, 0.7, _lib.PAIR_REFERENCE)
    monkeypatch.setenv("HULO_PAIR_BATCH_ROWS", "900")         # forces several batches
    db = gpu.db(rows, off)
    try:
        po, oi, oj = gpu.match_pairs(db, pairs, 0.7, cap=8)    # too small first: grows and retries
    finally:
        db.free()
    assert po.tolist() == wo and oi.tolist() == wi and oj.tolist() == wj


def test_track_pairs_feed_host_propagation(gpu, orc):
    """trackAKAZE = consecutive-frame pairs on the GPU + host track propagation (E.3)."""
    V = 6
    rows, off = synth.image_collection(V, 500, 35, overlap=0.7)
    pairs = [(f, f + 1) for f in range(V - 1)]
    db = gpu.db(rows, off)
    try:
        po, oi, oj = gpu.match_pairs(db, pairs, 0.7)
    finally:
        db.free()
    feat = [int(off[f + 1] - off[f]) for f in range(V - 1)]
    f, t, i, j = orc.track_propagate(V, 4, feat, po.astype(np.int64), oi, oj)
    assert len(f) > 50 and (t - f >= 2).all() and (t - f < 4).all()


def test_sharded_candidates_merge(gpu, orc):
    """Row-sharded search on one GPU: each shard's local top-2 with global indices, merged by
    hulo_merge_top2, equals the single-table search (what the NCCL all-gather path computes)."""
    A, B, _ = synth.descriptor_sets(1500, 40000, 61)
    B[30000:30050] = B[100:150]                 # duplicates across shards: ties on (distance, index)
    dA = gpu.db(A)
    bounds = [0, 9000, 9001, 25000, 40000]
    cand = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        dB = gpu.db(B[lo:hi])
        idx, dist = gpu.knn2_sharded(dA, dB, lo)
        dB.free()
        cand.append(np.stack([dist[:, 0], idx[:, 0], dist[:, 1], idx[:, 1]], axis=1))
    dA.free()
    mi, md = gpu.merge_top2(np.stack(cand))
    ri, rd = orc.knn2(A, B)
    assert np.array_equal(mi, ri) and np.array_equal(md, rd)
