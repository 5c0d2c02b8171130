#!/usr/bin/env python
"""BASELINE.json config 1 (one query image, 2000 descriptors, against a 200k-descriptor map of 100
views) with the VIEWS sharded over the ranks: hulo_engine_localize_sharded.  Every rank holds the
engine, matches its range of views, the surviving matches are all-gathered, every rank finishes the
query.  A latency measurement: wall clock per query, max over ranks, median over the repetitions.
Launch: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_localize_views_sharded.py"""
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
    reps = 30
    g = HuloGpu(int(os.environ.get("LOCAL_RANK", "0")))
    path = None
    if world > 1:
        uid, path = bench.rendezvous_id(rank, world, HuloGpu.comm_unique_id)
        g.comm_init(uid, rank, world)
    sc = synth.localization_scene(100, 2000, 20000, 2000, 1000)
    eng = LocalizeEngine(g, sc["rows"], sc["seg_offsets"], sc["obs_view"], sc["obs_feat"], sc["obs_landmark"],
                         sc["landmark_X"], sc["K"], ratio=0.6)
    for k in range(5):
        r = eng.localize_sharded(sc["q_desc"], sc["q_xy"], seed=k)
    wall, stages = [], []
    for k in range(reps):
        if world > 1:
            g.comm_barrier()
        t0 = time.perf_counter()
        r = eng.localize_sharded(sc["q_desc"], sc["q_xy"], seed=100 + k)
        dt = (time.perf_counter() - t0) * 1e3
        wall.append(g.comm_max(dt) if world > 1 else dt)
        stages.append(r["times_ms"])
    if rank == 0:
        st = np.median(np.array(stages), axis=0)
        print(json.dumps({"config": "C1 one query 2000 x 200000 (100 views), views sharded over the ranks",
                          "n_gpus": world, "ms_per_query": float(np.median(wall)),
                          "stage_ms_rank0": {"putMatch_incl_exchange": float(st[0]), "assembly": float(st[1]), "PnP": float(st[2])},
                          "localized": bool(r["localized"]),
                          "centre_error_m": float(np.linalg.norm(r["center"] - sc["center"])),
                          "collective": "one all-gather of the surviving matches (12 bytes each) per query"}), flush=True)
    eng.close()
    if world > 1:
        g.comm_barrier()
    g.close()
    if rank == 0 and path and os.path.exists(path):
        os.remove(path)


if __name__ == "__main__":
    main()
