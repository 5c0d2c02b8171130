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


def main():
    mode = sys.argv[1]
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
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
