"""GPU parity suite specific to K1t (the tensor-core engine of the flat 2-NN search, knn2_tc.cu):
tile-image bookkeeping, chunk / tile boundaries, tie order across tiles and chunks.  The generic
K1 expectations run on both engines in test_knn2_gpu.py."""
import numpy as np
import pytest

from sfmlocalization_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def tc_engine(gpu):
    gpu.set_knn_engine("tc")
    assert gpu.knn_engine == "tc"
    yield
    gpu.set_knn_engine("auto")


@pytest.mark.parametrize("nA,nB", [(127, 255), (128, 256), (129, 257), (1, 100000), (385, 65537), (7000, 2300),
                                   (20000, 3000), (640, 300001)])
def test_tile_and_chunk_boundaries(gpu, orc, nA, nB):
    """Searcher tiles of 128 rows, accumulator tiles of 256 rows, chunks a multiple of 256 rows:
    sizes on and next to every boundary."""
    A, B, _ = synth.descriptor_sets(nA, nB, 7 + nA + nB)
    idx, dist = gpu.knn2_host(A, B)
    ri, rd = orc.knn2(A, B)
    assert np.array_equal(dist, rd)
    assert np.array_equal(idx, ri)


def test_duplicates_across_tiles_and_chunks(gpu, orc):
    """The same row planted in several accumulator tiles and chunks: the lowest index must be the
    first neighbour and the next lowest the second, whatever tile or chunk holds them."""
    rng = np.random.default_rng(5)
    B = rng.integers(0, 256, size=(200000, 64), dtype=np.uint8)
    A = rng.integers(0, 256, size=(300, 64), dtype=np.uint8)
    spots = [255, 256, 70000, 70001, 199999]
    for k, a in enumerate(range(0, 300, 30)):
        for s in spots[k % 3:]:
            B[s - k] = A[a]
    idx, dist = gpu.knn2_host(A, B)
    ri, rd = orc.knn2(A, B)
    assert np.array_equal(dist, rd) and np.array_equal(idx, ri)
    assert (dist[::30, :] == 0).all()


def test_image_follows_table_updates(gpu, orc):
    """The tile image of a resident table is rebuilt after hulo_db_update (fewer rows included)."""
    A, B, _ = synth.descriptor_sets(700, 40000, 11)
    A2, B2, _ = synth.descriptor_sets(333, 25000, 12)
    dA, dB = gpu.db(A), gpu.db(B)
    try:
        i1, d1 = gpu.knn2(dA, dB)
        i1b, d1b = gpu.knn2(dA, dB)                  # cached images
        dB.update(B2)
        i2, d2 = gpu.knn2(dA, dB)
        dA.update(A2)
        i3, d3 = gpu.knn2(dA, dB)
    finally:
        dA.free(); dB.free()
    r1 = orc.knn2(A, B); r2 = orc.knn2(A, B2); r3 = orc.knn2(A2, B2)
    assert np.array_equal(i1, r1[0]) and np.array_equal(d1, r1[1])
    assert np.array_equal(i1b, r1[0]) and np.array_equal(d1b, r1[1])
    assert np.array_equal(i2, r2[0]) and np.array_equal(d2, r2[1])
    assert np.array_equal(i3, r3[0]) and np.array_equal(d3, r3[1])


def test_engines_agree_at_scale(gpu):
    """4096 x 1.5M: the two engines must return identical arrays (no oracle at this size)."""
    nA, nB = 4096, 1_500_000
    B = synth.random_rows(nB, 21)
    A = synth.random_rows(nA, 22)
    A, target = synth.plant_matches(A, B, 23, frac=0.3)
    dA, dB = gpu.db(A), gpu.db(B)
    try:
        it, dt = gpu.knn2(dA, dB)
        gpu.set_knn_engine("int")
        ii, di = gpu.knn2(dA, dB)
    finally:
        dA.free(); dB.free()
    assert np.array_equal(it, ii) and np.array_equal(dt, di)
    hit = target >= 0
    assert np.array_equal(it[hit, 0], target[hit])


def test_random_shapes_all_engines_agree(gpu):
    """Random table sizes (not aligned to any tile, group or chunk size): the three engines return
    identical arrays.  The integer engine is the one checked against the oracle everywhere else."""
    rng = np.random.default_rng(77)
    for trial in range(10):
        nA = int(rng.integers(1, 3000)); nB = int(rng.integers(1, 300000))
        A, B, _ = synth.descriptor_sets(nA, nB, 1000 + trial)
        gpu.set_knn_engine("int")
        want = gpu.knn2_host(A, B)
        for eng in ("tc", "tc8"):
            gpu.set_knn_engine(eng)
            got = gpu.knn2_host(A, B)
            assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), (eng, nA, nB)


def test_ragged_segments_item_mode(gpu, orc):
    """Views of every size from empty to a few tiles (segments aligned to nothing), a query that is
    not a multiple of the tile width: the item mode of both tensor-core forms against the oracle."""
    rng = np.random.default_rng(5)
    sizes = [0, 1, 2, 7, 8, 9, 127, 128, 129, 191, 192, 193, 223, 224, 225, 255, 256, 257, 700, 1031]
    rows = synth.random_rows(int(np.sum(sizes)), 31)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    query, _ = synth.plant_matches(synth.random_rows(777, 32), rows, 33, frac=0.6)
    db = gpu.db(rows, off)
    try:
        for eng in ("tc", "tc8"):
            gpu.set_knn_engine(eng)
            m = gpu.match_to_query(db, query, 0.8)
            k = 0
            for v, n_v in enumerate(sizes):
                a = rows[int(off[v]):int(off[v + 1])]
                oi, oj, od = orc.match_view_to_query(a, query, 0.8)
                n = len(oi)
                assert m["view_counts"][v] == n, (eng, v)
                assert np.array_equal(m["i"][k:k + n], oi) and np.array_equal(m["j"][k:k + n], oj)
                assert np.array_equal(m["d0"][k:k + n], od)
                k += n
            assert k == len(m["i"]) and k > 100
            pairs = [(18, 19), (19, 18), (5, 18), (17, 16), (3, 19), (0, 18)]
            o, pi, pj = gpu.match_pairs(db, pairs, 0.8)
            for p, (I, J) in enumerate(pairs):
                wi, wj = orc.match_pair(rows[int(off[I]):int(off[I + 1])], rows[int(off[J]):int(off[J + 1])], 0.8)
                assert np.array_equal(pi[int(o[p]):int(o[p + 1])], wi) and np.array_equal(pj[int(o[p]):int(o[p + 1])], wj), (eng, I, J)
    finally:
        db.free()
