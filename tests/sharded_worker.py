"""One rank of the row-sharded 2-NN check.  Launched by the tests with RANK / WORLD_SIZE /
LOCAL_RANK / MASTER_PORT in the environment (the same variables torchrun sets).
  mode "nccl": GPU path, hulo_knn2_sharded over an NCCL communicator (one process per GPU)
  mode "gloo": CPU restatement of the same exchange (oracle top-2 per shard, all_gather over
               torch.distributed gloo, lexicographic merge) -- covers the host-side sharding
               logic where there is no GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sfmlocalization_b200 import synth  # noqa: E402

NA, NB = 700, 30011


def shard_bounds(n, world):
    per = (n + world - 1) // world
    return [(min(r * per, n), min((r + 1) * per, n)) for r in range(world)]


def merge_candidates(cands):
    """cands: world x nA x 4 {d0, i0, d1, i1}; -1 index = missing."""
    d = np.concatenate([cands[:, :, 0], cands[:, :, 2]], axis=0).T.astype(np.int64)
    i = np.concatenate([cands[:, :, 1], cands[:, :, 3]], axis=0).T.astype(np.int64)
    d = np.where(i < 0, np.iinfo(np.int64).max, d)
    order = np.lexsort((i, d), axis=1)[:, :2]
    return np.take_along_axis(i, order, axis=1).astype(np.int32), np.take_along_axis(d, order, axis=1).astype(np.int32)


def localize_views_gloo(rank, world):
    """View-sharded query (hulo_engine_localize_sharded) on the CPU: the library's own partition of
    the view list, the CPU restatement's matcher on this rank's views, an all-gather of the padded
    blocks {n, counts, (i, j, d0)...} over gloo, concatenation in rank order -- must equal the
    unsharded match list, view by view."""
    import torch
    import torch.distributed as dist_
    from oracle import oracle as orc
    from sfmlocalization_b200.gpu import partition_views
    dist_.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["MASTER_PORT"],
                             rank=rank, world_size=world)
    sc = synth.localization_scene(13, 300, 900, 400, 31)
    off = sc["seg_offsets"].astype(np.int64)
    views = [11, 2, 3, 4, 5, 12, 0, 7]                      # a selection, not in id order
    rows = [off[v + 1] - off[v] for v in views]
    bounds = partition_views(rows, world)
    mine = views[bounds[rank]:bounds[rank + 1]]
    max_nv = int(np.max(np.diff(bounds)))

    def match(vs):
        out = []
        for v in vs:
            oi, oj, od = orc.match_view_to_query(sc["rows"][off[v]:off[v + 1]], sc["q_desc"], 0.6)
            out.append(np.stack([oi, oj, od], axis=1).astype(np.int64).reshape(-1, 3))
        return out

    local = match(mine)
    n_m = sum(len(x) for x in local)
    slots = 4096
    block = np.zeros(1 + max_nv + 3 * slots, np.int64)
    block[0] = n_m
    block[1:1 + len(local)] = [len(x) for x in local]
    if n_m:
        block[1 + max_nv:1 + max_nv + 3 * n_m] = np.concatenate(local).ravel()
    parts = [torch.empty(len(block), dtype=torch.int64) for _ in range(world)]
    dist_.all_gather(parts, torch.from_numpy(block))
    dist_.destroy_process_group()
    counts, recs = [], []
    for r in range(world):
        b = parts[r].numpy()
        counts += b[1:1 + (bounds[r + 1] - bounds[r])].tolist()
        recs.append(b[1 + max_nv:1 + max_nv + 3 * int(b[0])].reshape(-1, 3))
    recs = np.concatenate(recs)
    want = match(views)
    ok = counts == [len(x) for x in want] and np.array_equal(recs, np.concatenate(want))
    ok = ok and bounds[0] == 0 and bounds[-1] == len(views) and sum(counts) > 100
    print("rank %d/%d views-gloo: %s" % (rank, world, "OK" if ok else "MISMATCH"))
    sys.exit(0 if ok else 1)


def localize_views_nccl(rank, world):
    """hulo_engine_localize_sharded on `world` GPUs: every rank gets the single-GPU answer, bit for bit."""
    import bench
    from sfmlocalization_b200.gpu import HuloGpu, LocalizeEngine
    g = HuloGpu(int(os.environ.get("LOCAL_RANK", rank)))
    uid, path = bench.rendezvous_id(rank, world, HuloGpu.comm_unique_id)
    sc = synth.localization_scene(24, 800, 4000, 900, 1)
    eng = LocalizeEngine(g, sc["rows"], sc["seg_offsets"], sc["obs_view"], sc["obs_feat"], sc["obs_landmark"],
                         sc["landmark_X"], sc["K"], ratio=0.6)
    want = [eng.localize(sc["q_desc"], sc["q_xy"], seed=5),
            eng.localize(sc["q_desc"], sc["q_xy"], views=[9, 2, 3, 4, 11, 17, 20], seed=6)]
    eng.set_keypoints(sc["map_xy"], sc["view_wh"], synth.IMAGE_WH)
    eng.configure_geometric(True, 25, 4.0)
    eng.set_guided_matching(True)
    want.append(eng.localize(sc["q_desc"], sc["q_xy"], seed=7))
    eng.configure_geometric(False)
    g.comm_init(uid, rank, world)
    got = [eng.localize_sharded(sc["q_desc"], sc["q_xy"], seed=5),
           eng.localize_sharded(sc["q_desc"], sc["q_xy"], views=[9, 2, 3, 4, 11, 17, 20], seed=6)]
    eng.configure_geometric(True, 25, 4.0)
    eng.set_guided_matching(True)
    got.append(eng.localize_sharded(sc["q_desc"], sc["q_xy"], seed=7))
    one = eng.localize_sharded(sc["q_desc"], sc["q_xy"], views=[3], seed=8)     # fewer views than ranks
    # the same query with the matching forced onto the tensor cores (item mode) and onto the integer pipes
    eng.configure_geometric(False)
    per_engine = []
    for engine in ("tc", "int"):
        g.set_knn_engine(engine)
        per_engine.append(eng.localize_sharded(sc["q_desc"], sc["q_xy"], seed=5))
    g.set_knn_engine("auto")
    got += per_engine
    want += [want[0], want[0]]
    g.comm_barrier()
    ok = all(w["localized"] and a["localized"] and np.array_equal(a["center"], w["center"]) and
             np.array_equal(a["R"], w["R"]) and np.array_equal(a["corr_qfeat"], w["corr_qfeat"]) and
             np.array_equal(a["corr_landmark"], w["corr_landmark"]) and np.array_equal(a["inliers"], w["inliers"])
             for a, w in zip(got, want))
    ok = ok and len(one["corr_qfeat"]) > 0
    eng.close(); g.close()
    if rank == 0 and os.path.exists(path):
        os.remove(path)
    print("rank %d/%d views-nccl: %s" % (rank, world, "OK" if ok else "MISMATCH"))
    sys.exit(0 if ok else 1)


def main():
    mode = sys.argv[1]
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    if mode == "views-gloo":
        return localize_views_gloo(rank, world)
    if mode == "views-nccl":
        return localize_views_nccl(rank, world)
    A, B, _ = synth.descriptor_sets(NA, NB, 123)
    B[20000:20040] = B[5:45]                        # cross-shard duplicates
    lo, hi = shard_bounds(NB, world)[rank]
    from oracle import oracle as orc
    want_i, want_d = orc.knn2(A, B)
    if mode == "nccl":
        sys.path.insert(0, ROOT)
        import bench
        from sfmlocalization_b200.gpu import HuloGpu
        g = HuloGpu(int(os.environ.get("LOCAL_RANK", rank)))
        uid, path = bench.rendezvous_id(rank, world, HuloGpu.comm_unique_id)
        g.comm_init(uid, rank, world)
        dA, dB = g.db(A), g.db(B[lo:hi])
        # both exchange flavours: stores into peer-mapped buffers (default) and ncclAllGather
        results = []
        for flavour in ("peer", "nccl", "peer"):
            os.environ["HULO_EXCHANGE"] = flavour
            for _ in range(3):                       # repeated calls exercise the parity double buffer
                results.append(g.knn2_sharded(dA, dB, lo))
        # a larger searcher table forces the exchange buffers to be rebuilt (collective)
        os.environ["HULO_EXCHANGE"] = "peer"
        A2 = np.concatenate([A] * 8, axis=0)
        dA2 = g.db(A2)
        i2, d2 = g.knn2_sharded(dA2, dB, lo)
        dA2.free()
        # every arithmetic engine behind the same exchange: identical arrays
        for eng in ("int", "tc", "tc8"):
            g.set_knn_engine(eng)
            results.append(g.knn2_sharded(dA, dB, lo))
        g.set_knn_engine("auto")
        # pipelined calls (the exchange of one call overlaps K1 of the next): un-fetched searches
        # with two alternating searcher tables, the result read afterwards must be the last
        # call's; then an unsharded search issued right behind a pending exchange
        A_rev = np.ascontiguousarray(A[::-1])
        dAr = g.db(A_rev)
        for eng in ("tc", "int"):
            g.set_knn_engine(eng)
            for k in range(7):
                g.knn2_sharded(dAr if k % 2 else dA, dB, lo, fetch=False)
            pi, pd = g.knn2_fetch(NA)
            assert np.array_equal(pi, results[0][0]) and np.array_equal(pd, results[0][1])
            g.knn2_sharded(dAr, dB, lo, fetch=False)
            pi, pd = g.knn2_fetch(NA)
            assert np.array_equal(pi, results[0][0][::-1]) and np.array_equal(pd, results[0][1][::-1])
            g.knn2_sharded(dA, dB, lo, fetch=False)
            li_, ld_ = g.knn2(dAr, dB)                # local shard only, behind the pending exchange
            wi_, wd_ = orc.knn2(A_rev, B[lo:hi])
            assert np.array_equal(li_, wi_) and np.array_equal(ld_, wd_)
        g.set_knn_engine("auto")
        # submit / collect: results one search behind, each equal to the blocking call
        g.knn2_sharded_submit(dA, dB, lo)
        g.knn2_sharded_submit(dAr, dB, lo)
        first = g.knn2_sharded_collect()
        g.knn2_sharded_submit(dA, dB, lo)
        second = g.knn2_sharded_collect()
        third = g.knn2_sharded_collect()
        assert np.array_equal(first[0], results[0][0]) and np.array_equal(first[1], results[0][1])
        assert np.array_equal(second[0], results[0][0][::-1]) and np.array_equal(second[1], results[0][1][::-1])
        assert np.array_equal(third[0], results[0][0]) and np.array_equal(third[1], results[0][1])
        dAr.free()
        for (i_, d_) in results:
            assert np.array_equal(i_, results[0][0]) and np.array_equal(d_, results[0][1])
        assert np.array_equal(i2[:NA], results[0][0]) and np.array_equal(i2[-NA:], results[0][0])
        assert np.array_equal(d2[:NA], results[0][1])
        idx, dist = results[0]
        t = g.comm_max(float(rank))
        assert t == float(world - 1)
        g.comm_barrier()
        dA.free(); dB.free(); g.close()
        if rank == 0 and os.path.exists(path):
            os.remove(path)
    else:
        import torch
        import torch.distributed as dist_
        dist_.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["MASTER_PORT"],
                                 rank=rank, world_size=world)
        li, ld = orc.knn2(A, B[lo:hi])
        li = np.where(li >= 0, li + lo, -1)
        mine = torch.from_numpy(np.stack([ld[:, 0], li[:, 0], ld[:, 1], li[:, 1]], axis=1).astype(np.int32))
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist_.all_gather(parts, mine)
        idx, dist = merge_candidates(np.stack([p.numpy() for p in parts]))
        dist_.destroy_process_group()
    ok = np.array_equal(idx, want_i) and np.array_equal(dist, want_d)
    print("rank %d/%d %s: %s" % (rank, world, mode, "OK" if ok else "MISMATCH"))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
