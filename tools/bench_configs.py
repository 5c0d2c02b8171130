#!/usr/bin/env python
"""Secondary measurements on one B200 for the other BASELINE.json configs (bench.py carries the
headline C3 line).  Prints one JSON object per config; run under gpurun and keep the output in
profiles/.  Timing: CUDA events on the library stream (hulo_timer_*), 1 warm-up + N timed steps."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sfmlocalization_b200 import synth  # noqa: E402
from sfmlocalization_b200.gpu import HuloGpu  # noqa: E402


def timed(g, fn, steps):
    fn()
    g.synchronize()
    g.timer_start()
    for _ in range(steps):
        fn()
    return g.timer_stop() / steps


def flat(g, name, nA, nB, seed, steps=3):
    A, B, _ = synth.descriptor_sets(nA, nB, seed)
    dA, dB = g.db(A), g.db(B)
    ms = timed(g, lambda: g.knn2(dA, dB, fetch=False), steps)
    dA.free(); dB.free()
    print(json.dumps({"config": name, "nA": nA, "nB": nB, "ms": ms, "gdist_per_s": nA * nB / ms / 1e6}), flush=True)


def pairs(g, name, n_img, rows, n_pairs, seed):
    allrows, off = synth.image_collection(n_img, rows, seed, overlap=0.3)
    db = g.db(allrows, off)
    pl = [(a, b) for a in range(n_img) for b in range(a + 1, n_img)][:n_pairs]
    g.match_pairs(db, pl, 0.7, cap=len(pl) * rows)           # warm-up at full size: scratch buffers grow here
    dts = []
    for _ in range(3):
        t0 = time.perf_counter()
        off_, oi, oj = g.match_pairs(db, pl, 0.7, cap=len(pl) * rows)
        dts.append(time.perf_counter() - t0)
    dt = float(np.median(dts))
    db.free()
    dist = len(pl) * rows * rows
    print(json.dumps({"config": name, "images": n_img, "rows": rows, "pairs": len(pl), "wall_ms": dt * 1e3,
                      "gdist_per_s_wall": dist / dt / 1e9, "matches": int(len(oi)),
                      "note": "wall clock of hulo_match_pairs incl. item list upload, post filters, D2H"}), flush=True)


def scoring(g, name, H, N, seed):
    sc = synth.resection_scene(N, seed, outlier_frac=0.5)
    tri = synth.sample_triplets(N, H // 4, seed)
    models, nm = g.p3p(tri, sc["x2d"], sc["X3d"], sc["K"])
    models = models.reshape(-1, 3, 4)
    g.score_resection(models[:64], sc["x2d"], sc["X3d"], sc["K"])
    t0 = time.perf_counter()
    g.score_resection(models, sc["x2d"], sc["X3d"], sc["K"])
    dt = time.perf_counter() - t0
    print(json.dumps({"config": name, "hypotheses": int(models.shape[0]), "N": N, "wall_ms": dt * 1e3,
                      "hyp_per_s": models.shape[0] / dt,
                      "note": "wall clock of hulo_score_resection incl. H2D of models and D2H of scores"}), flush=True)


def resection(g, name, N, outl, seed):
    """AC-RANSAC resection (4096 iterations) in both schedules: batched (engine default) and sequential."""
    sc = synth.resection_scene(N, seed, outlier_frac=outl)
    out = {"config": name, "N": N, "outlier_frac": outl}
    for label, seq in (("batched", False), ("sequential", True)):
        g.resect_acransac(sc["x2d"], sc["X3d"], sc["K"], seed=1, sequential=seq)
        ts = []
        for k in range(10):
            t0 = time.perf_counter()
            r = g.resect_acransac(sc["x2d"], sc["X3d"], sc["K"], seed=10 + k, sequential=seq)
            ts.append(time.perf_counter() - t0)
        out[label + "_ms"] = float(np.median(ts)) * 1e3
        out[label + "_inliers"] = int(len(r["inliers"]))
    print(json.dumps(out), flush=True)


def batched_server(g, name, n_views, feats, n_landmarks, n_queries, nq, seed):
    """BASELINE.json config 4 end to end: concurrent query images against one resident map."""
    from sfmlocalization_b200.gpu import LocalizeEngine
    t0 = time.perf_counter()
    sc = synth.localization_scene(n_views, feats, n_landmarks, nq, seed, window=6000)
    qs = [synth.extra_query(sc, nq, seed + 10 + k) for k in range(n_queries)]
    gen_s = time.perf_counter() - t0
    eng = LocalizeEngine(g, sc["rows"], sc["seg_offsets"], sc["obs_view"], sc["obs_feat"], sc["obs_landmark"],
                         sc["landmark_X"], sc["K"], ratio=0.6)
    descs = [q["q_desc"] for q in qs]; xys = [q["q_xy"] for q in qs]
    eng.localize_batch(descs[:2], xys[:2])
    t0 = time.perf_counter()
    b = eng.localize_batch(descs, xys, seed=5)
    dt = time.perf_counter() - t0
    eng.close()
    err = [float(np.linalg.norm(b["center"][k] - qs[k]["center"])) for k in range(n_queries) if b["localized"][k]]
    dist = float(n_queries) * nq * sc["rows"].shape[0]
    print(json.dumps({"config": name, "queries": n_queries, "query_rows": nq, "map_rows": int(sc["rows"].shape[0]),
                      "wall_s": dt, "localizations_per_s": n_queries / dt, "gdist_per_s_wall": dist / dt / 1e9,
                      "stage_ms_total": {"putMatch": b["times_ms"][0], "assembly": b["times_ms"][1], "PnP": b["times_ms"][2]},
                      "fraction_localized": float(b["localized"].mean()),
                      "centre_error_m_median": float(np.median(err)) if err else None,
                      "scene_generation_s": gen_s}), flush=True)


def main():
    which = sys.argv[1:] or ["c1", "c1r", "c2", "c4", "c5", "k2"]
    with HuloGpu(0) as g:
        if "c1" in which:
            flat(g, "C1 query->map 2000 x 200000", 2000, 200000, 1000, steps=20)
        if "c1r" in which:
            flat(g, "C1 reference direction 200000 x 2000", 200000, 2000, 1001, steps=20)
        if "c2" in which:
            pairs(g, "C2 pairwise 200 x 5000 (first 2000 of 19900 pairs)", 200, 5000, 2000, 2000)
        if "small" in which:
            flat(g, "small searcher set 500 x 200000", 500, 200000, 1002, steps=20)
            flat(g, "mid searcher set 1000 x 200000", 1000, 200000, 1003, steps=20)
        if "c4" in which:
            flat(g, "C4 768000 x 2000000", 768000, 2000000, 4000, steps=1)
        if "c5" in which:
            flat(g, "C5 one of 8 shards: 16384 x 6250000", 16384, 6250000, 5000, steps=2)
        if "c4e" in which:
            batched_server(g, "C4 batched server: 256 queries x 3000 vs 2M-descriptor map, end to end", 1000, 2000,
                           200000, 256, 3000, 4100)
        if "pnp" in which:
            resection(g, "AC-RANSAC resection, clean query (C1-like)", 700, 0.02, 4200)
            resection(g, "AC-RANSAC resection, half outliers", 700, 0.5, 4201)
            resection(g, "AC-RANSAC resection, 2000 correspondences, 70 % outliers", 2000, 0.7, 4202)
        if "k2" in which:
            for N in (100, 500, 2000):
                scoring(g, "K2 scoring 4096 triplets x <=4 models", 16384, N, 4000 + N)


if __name__ == "__main__":
    main()
