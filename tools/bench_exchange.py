#!/usr/bin/env python
"""Latency of the row-sharded exchange step on its own: a tiny shard (so K1 is negligible) and
4096 / 16384 / 262144 searcher rows, fused peer-store exchange vs ncclAllGather.
Launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_exchange.py
Prints one JSON line per (rows, flavour) on rank 0; time = CUDA events, max over ranks."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sfmlocalization_b200 import synth  # noqa: E402
from sfmlocalization_b200.gpu import HuloGpu  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    g = HuloGpu(int(os.environ.get("LOCAL_RANK", "0")))
    uid, path = bench.rendezvous_id(rank, world, HuloGpu.comm_unique_id)
    g.comm_init(uid, rank, world)
    shard = synth.random_rows(256, 10 + rank)
    dB = g.db(shard)
    for nA in (4096, 16384, 262144):
        dA = g.db(synth.random_rows(nA, 1))
        for flavour in ("peer", "nccl"):
            os.environ["HULO_EXCHANGE"] = flavour
            for _ in range(5):
                g.knn2_sharded(dA, dB, rank * 256, fetch=False)
            g.comm_barrier()
            steps = 200
            g.timer_start()
            for _ in range(steps):
                g.knn2_sharded(dA, dB, rank * 256, fetch=False)
            ms = g.comm_max(g.timer_stop()) / steps
            if rank == 0:
                print(json.dumps({"rows": nA, "world": world, "exchange": flavour, "us_per_step": ms * 1e3,
                                  "bytes_per_rank": nA * 16,
                                  "note": "K1 over a 256-row shard + chunk merge + exchange + rank merge"}), flush=True)
        dA.free()
    dB.free()
    g.comm_barrier()
    g.close()
    if rank == 0 and os.path.exists(path):
        os.remove(path)


if __name__ == "__main__":
    main()
