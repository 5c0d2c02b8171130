#!/usr/bin/env python
"""BASELINE.json config 4 across GPUs: 256 concurrent query images x 3000 descriptors against a
2M-descriptor map, end to end (hulo_engine_localize_batch: batched matching, view filter, 2D-3D
assembly, AC-RANSAC resection, pose).  The queries are independent units: every rank holds the
whole map (128 MB) and localises its share of the queries -- replicas, NO data-path collective.
Launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_localize_sharded.py
Rank 0 prints one JSON line; time = wall clock of the batched call, max over ranks."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sfmlocalization_b200 import synth  # noqa: E402
from sfmlocalization_b200.gpu import HuloGpu, LocalizeEngine  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    n_queries, nq, seed = 256, 3000, 4100
    g = HuloGpu(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        uid, path = bench.rendezvous_id(rank, world, HuloGpu.comm_unique_id)
        g.comm_init(uid, rank, world)
    sc = synth.localization_scene(1000, 2000, 200000, nq, seed, window=6000)
    mine = list(range(rank, n_queries, world))
    qs = [synth.extra_query(sc, nq, seed + 10 + k) for k in mine]
    eng = LocalizeEngine(g, sc["rows"], sc["seg_offsets"], sc["obs_view"], sc["obs_feat"], sc["obs_landmark"],
                         sc["landmark_X"], sc["K"], ratio=0.6)
    descs = [q["q_desc"] for q in qs]; xys = [q["q_xy"] for q in qs]
    eng.localize_batch(descs[:2], xys[:2])
    if world > 1:
        g.comm_barrier()
    t0 = time.perf_counter()
    b = eng.localize_batch(descs, xys, seed=5)
    dt = time.perf_counter() - t0
    dt = g.comm_max(dt) if world > 1 else dt
    ok = float(b["localized"].sum())
    err = [float(np.linalg.norm(b["center"][k] - qs[k]["center"])) for k in range(len(qs)) if b["localized"][k]]
    worst = g.comm_max(max(err) if err else 0.0) if world > 1 else (max(err) if err else 0.0)
    n_ok = -g.comm_max(-ok) if world > 1 else ok          # min over ranks of the localised count
    if rank == 0:
        print(json.dumps({"config": "C4 batched server: 256 queries x 3000 vs 2M-descriptor map, end to end",
                          "n_gpus": world, "queries_per_rank": len(mine), "wall_s": dt,
                          "localizations_per_s": n_queries / dt, "min_localized_on_a_rank": int(n_ok),
                          "worst_centre_error_m": worst, "collective": "none (barrier + max for timing only)",
                          "sharding": "queries over ranks, map replicated"}), flush=True)
    eng.close()
    if world > 1:
        g.comm_barrier()
    g.close()
    if rank == 0 and world > 1 and os.path.exists(path):
        os.remove(path)


if __name__ == "__main__":
    main()
