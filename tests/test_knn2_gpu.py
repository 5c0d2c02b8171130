"""GPU parity suite for K1 (exact Hamming 2-NN) through the C-ABI: bit-exact against the CPU
oracle and the committed OpenCV golden vectors."""
import numpy as np
import pytest

from sfmlocalization_b200 import synth

pytestmark = pytest.mark.gpu
INT_MAX = 2**31 - 1


@pytest.fixture(autouse=True, params=["int", "tc", "tc8"])
def engine(request, gpu):
    """Every test of this module runs on both arithmetic engines of the flat search: the integer
    pipes (K1) and the int8 tensor-core contraction (K1t).  Same bit-exact expectations."""
    gpu.set_knn_engine(request.param)
    yield request.param
    gpu.set_knn_engine("auto")


def run_gpu(gpu, A, B):
    return gpu.knn2_host(A, B)


@pytest.mark.parametrize("name", ["tie", "pl", "w61", "dup"])
def test_golden_vectors(gpu, golden, name):
    idx, dist = run_gpu(gpu, golden[name + "_A"], golden[name + "_B"])
    assert np.array_equal(idx, golden[name + "_idx"])
    assert np.array_equal(dist, golden[name + "_dist"])


@pytest.mark.parametrize("nA,nB", [(1, 1), (1, 2), (5, 0), (5, 1), (5, 2), (5, 3), (0, 10), (33, 65), (257, 63),
                                   (2049, 129), (300, 4097), (4096, 1000), (1000, 20000)])
def test_shapes_vs_oracle(gpu, orc, nA, nB):
    A, B, _ = synth.descriptor_sets(nA, nB, 100 + nA + nB)
    idx, dist = run_gpu(gpu, A, B)
    ri, rd = orc.knn2(A, B)
    assert np.array_equal(dist, rd)
    assert np.array_equal(idx, ri)


def test_tie_heavy_large(gpu, orc):
    """Distances tie constantly; every merge level must keep (distance, index) order."""
    B = synth.tie_heavy_rows(50000, 7, varying_bits=10)
    A = synth.tie_heavy_rows(3000, 8, varying_bits=10)
    idx, dist = run_gpu(gpu, A, B)
    ri, rd = orc.knn2(A, B)
    assert np.array_equal(dist, rd) and np.array_equal(idx, ri)
    assert (dist[:, 0] == dist[:, 1]).mean() > 0.5


def test_all_identical_rows(gpu):
    B = np.tile(synth.random_rows(1, 3), (5000, 1))
    A = B[:100].copy()
    idx, dist = run_gpu(gpu, A, B)
    assert (dist == 0).all() and (idx[:, 0] == 0).all() and (idx[:, 1] == 1).all()


def test_full_distance_512(gpu):
    """Rows that differ in all 512 bits (not producible by AKAZE, but the key packing must hold)."""
    B = np.zeros((10, 64), np.uint8)
    A = np.full((3, 64), 0xFF, np.uint8)
    idx, dist = run_gpu(gpu, A, B)
    assert (dist == 512).all() and (idx[:, 0] == 0).all() and (idx[:, 1] == 1).all()


def test_device_resident_tables(gpu, orc):
    A, B, target = synth.descriptor_sets(2500, 30000, 77)
    dA, dB = gpu.db(A), gpu.db(B)
    try:
        idx, dist = gpu.knn2(dA, dB)
        assert gpu.knn2(dA, dB, fetch=False) is None
        idx_b, dist_b = gpu.knn2_fetch(len(dA))
    finally:
        dA.free(); dB.free()
    ri, rd = orc.knn2(A, B)
    assert np.array_equal(idx, ri) and np.array_equal(dist, rd)
    assert np.array_equal(idx_b, ri) and np.array_equal(dist_b, rd)
    hit = target >= 0
    assert np.array_equal(idx[hit, 0], target[hit])


def test_submit_collect_pipeline(gpu, orc):
    """hulo_knn2_sharded_submit / _collect on a world of one: a stream of searcher batches through one
    table object, results collected one search behind, each equal to the blocking call; the
    outstanding-search limits are errors, not overwrites."""
    from sfmlocalization_b200.gpu import HuloError, check
    batches = [synth.descriptor_sets(900, 40000, 300 + k)[0] for k in range(5)]
    _, B, _ = synth.descriptor_sets(8, 40000, 299)
    want = [orc.knn2(a, B) for a in batches]
    dA, dB = gpu.db(batches[0]), gpu.db(B)
    try:
        got = []
        gpu.knn2_sharded_submit(dA, dB, 0)
        for k in range(1, len(batches)):
            dA.update(batches[k])
            gpu.knn2_sharded_submit(dA, dB, 0)
            got.append(gpu.knn2_sharded_collect())
        with pytest.raises(HuloError):                    # a plain search would clobber the outstanding one
            gpu.knn2(dA, dB)
        got.append(gpu.knn2_sharded_collect())
        for (gi, gd), (wi, wd) in zip(got, want):
            assert np.array_equal(gi, wi) and np.array_equal(gd, wd)
        with pytest.raises(HuloError):                    # nothing left to collect
            check(gpu.lib.hulo_knn2_sharded_collect(gpu.h, None, None, None))
        gpu.knn2_sharded_submit(dA, dB, 0)
        gpu.knn2_sharded_submit(dA, dB, 0)
        with pytest.raises(HuloError):                    # a third outstanding search
            gpu.knn2_sharded_submit(dA, dB, 0)
        a = gpu.knn2_sharded_collect(); b = gpu.knn2_sharded_collect()
        assert np.array_equal(a[0], want[-1][0]) and np.array_equal(b[1], want[-1][1])
        idx, dist = gpu.knn2(dA, dB)                      # plain calls work again
        assert np.array_equal(idx, want[-1][0]) and np.array_equal(dist, want[-1][1])
    finally:
        dA.free(); dB.free()


def test_planted_property_at_scale(gpu):
    """Full-size property check (no oracle): every planted row finds its source as first
    neighbour at the planted distance, and d0 <= d1, on a 4096 x 2M problem."""
    nA, nB = 4096, 2_000_000
    B = synth.random_rows(nB, 5)
    A = synth.random_rows(nA, 6)
    A, target = synth.plant_matches(A, B, 9, frac=0.3)
    idx, dist = run_gpu(gpu, A, B)
    hit = np.nonzero(target >= 0)[0]
    assert np.array_equal(idx[hit, 0], target[hit])
    x = np.bitwise_xor(A[hit], B[target[hit]])
    d_true = np.unpackbits(x, axis=1).sum(axis=1)
    assert np.array_equal(dist[hit, 0], d_true)
    assert (dist[:, 0] <= dist[:, 1]).all()
    assert (idx >= 0).all() and (idx < nB).all() and (idx[:, 0] != idx[:, 1]).all()
    # spot-check the reported second neighbours' distances
    rows = np.arange(0, nA, 97)
    x = np.bitwise_xor(A[rows], B[idx[rows, 1]])
    assert np.array_equal(np.unpackbits(x, axis=1).sum(axis=1), dist[rows, 1])


def test_split_database_merge_is_associative(gpu, orc):
    """Searching two halves and merging on (distance, global index) equals one search: the
    property the row-sharded multi-GPU mode relies on."""
    A, B, _ = synth.descriptor_sets(500, 9000, 55)
    i_full, d_full = run_gpu(gpu, A, B)
    cut = 4321
    i1, d1 = run_gpu(gpu, A, B[:cut])
    i2, d2 = run_gpu(gpu, A, B[cut:])
    i2 = i2 + cut
    cand_d = np.concatenate([d1, d2], axis=1).astype(np.int64)
    cand_i = np.concatenate([i1, i2], axis=1).astype(np.int64)
    order = np.lexsort((cand_i, cand_d), axis=1)[:, :2]
    md = np.take_along_axis(cand_d, order, axis=1); mi = np.take_along_axis(cand_i, order, axis=1)
    assert np.array_equal(md, d_full) and np.array_equal(mi, i_full)


@pytest.mark.parametrize("stride", [61, 64, 70])
def test_table_round_trip_through_the_device_layout(gpu, stride):
    """Rows are stored folded on the device (knn2.cuh); a download must give back the 64-byte
    zero-padded rows that were uploaded, for any stride, and an update must replace them."""
    rng = np.random.default_rng(stride)
    rows = rng.integers(0, 256, size=(1000, stride), dtype=np.uint8)
    want = np.zeros((1000, 64), np.uint8)
    want[:, :min(stride, 64)] = rows[:, :64]
    db = gpu.db(rows)
    try:
        assert np.array_equal(db.download(), want)
        assert np.array_equal(db.download(first=123, n=77), want[123:200])
        rows2 = rng.integers(0, 256, size=(640, stride), dtype=np.uint8)
        db.update(rows2)
        want2 = np.zeros((640, 64), np.uint8)
        want2[:, :min(stride, 64)] = rows2[:, :64]
        assert len(db) == 640 and np.array_equal(db.download(), want2)
    finally:
        db.free()


def test_full_512_bit_rows(gpu, orc):
    """Rows whose last bytes are not zero (every bit of every word live): exercises all six folded
    adder sums of the device layout, including the one over words 9..15."""
    rng = np.random.default_rng(99)
    A = rng.integers(0, 256, size=(777, 64), dtype=np.uint8)
    B = rng.integers(0, 256, size=(5003, 64), dtype=np.uint8)
    B[100] = A[5]; B[4000] = A[5]                     # an exact duplicate pair: distance 0 twice, lowest index first
    idx, dist = gpu.knn2_host(A, B)
    ri, rd = orc.knn2(A, B)
    assert np.array_equal(idx, ri) and np.array_equal(dist, rd)
    assert idx[5].tolist() == [100, 4000] and dist[5].tolist() == [0, 0]
