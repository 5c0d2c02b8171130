#!/usr/bin/env python
"""BASELINE.json config 2 across GPUs: exhaustive pairwise matching of 200 synthetic images x 5000
descriptors (19 900 pairs, ratio 0.7, one-to-one filter), the pair list partitioned over the ranks
(hulo::partitionPairs, LPT on n_I x n_J) with NO data-path collective: every rank holds all
descriptors (64 MB) and matches its own pairs (hulo_match_pairs); the ranks' outputs concatenate.
Launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_pairs_sharded.py
Rank 0 prints one JSON line; time = wall clock of the call, max over ranks (NCCL only for the
barrier and the max)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sfmlocalization_b200 import synth  # noqa: E402
from sfmlocalization_b200.gpu import HuloGpu  # noqa: E402
from tests import hostlib  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    n_img, rows = 200, 5000
    g = HuloGpu(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        uid, path = bench.rendezvous_id(rank, world, HuloGpu.comm_unique_id)
        g.comm_init(uid, rank, world)
    allrows, off = synth.image_collection(n_img, rows, 2000, overlap=0.3)
    db = g.db(allrows, off)
    pairs = np.array([(a, b) for a in range(n_img) for b in range(a + 1, n_img)], np.uint64)
    mine = hostlib.partition_pairs(pairs, np.full(n_img, rows), rank, world)
    pl = pairs[mine]
    g.match_pairs(db, pl, 0.7, cap=4 << 20)                      # warm-up at full size
    if world > 1:
        g.comm_barrier()
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        o, oi, oj = g.match_pairs(db, pl, 0.7, cap=4 << 20)
        dt = time.perf_counter() - t0
        times.append(g.comm_max(dt) if world > 1 else dt)
    dt = float(np.median(times))
    n_matches = g.comm_max(float(len(oi))) if world > 1 else len(oi)
    if rank == 0:
        dist = float(len(pairs)) * rows * rows
        print(json.dumps({"config": "C2 pairwise 200 x 5000, all 19900 pairs", "n_gpus": world, "pairs_per_rank": int(len(pl)),
                          "wall_s": dt, "gdist_per_s": dist / dt / 1e9, "pairs_per_s": len(pairs) / dt,
                          "max_matches_on_a_rank": int(n_matches), "collective": "none (barrier + max for timing only)",
                          "note": "wall clock of hulo_match_pairs on each rank's share incl. item upload, ratio / "
                                  "one-to-one filters, compaction and D2H; max over ranks"}), flush=True)
    db.free()
    if world > 1:
        g.comm_barrier()
    g.close()
    if rank == 0 and world > 1 and os.path.exists(path):
        os.remove(path)


if __name__ == "__main__":
    main()
