#!/usr/bin/env python
"""The matching stages of ExtFeatAndMatch (computeFeaturesAndMatches.cpp:150-246) end to end on one
GPU, through the C-ABI, on a synthetic collection WITH geometry: V views of one scene (each with its
own pose, F features of which 60 % observe landmarks from a sliding window), all V(V-1)/2 pairs:
  putative matching (ratio 0.6, one-to-one)  ->  drop pairs below minMatch 60  ->
  F-matrix AC-RANSAC (ransacRound 500, 4 px: ReconstructParam.py:76)  ->  guided matching (-gm, the
  reconstruction default, ReconstructParam.py:70-71).
Prints one JSON line with the wall time of every stage."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from sfmlocalization_b200 import synth  # noqa: E402
from sfmlocalization_b200.gpu import HuloGpu  # noqa: E402


def main():
    V = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    F = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
    t0 = time.perf_counter()
    sc = synth.localization_scene(V, F, 12 * F, 10, 77, track_frac=0.6, window=4 * F)
    gen = time.perf_counter() - t0
    off = sc["seg_offsets"]
    pairs = np.array([(a, b) for a in range(V) for b in range(a + 1, V)], np.uint32)
    w, h = synth.IMAGE_WH
    with HuloGpu(0) as g:
        db = g.db(sc["rows"], off)
        g.match_pairs(db, pairs, 0.6, cap=len(pairs) * F // 4)        # warm-up at full size
        t0 = time.perf_counter()
        po, pi, pj = g.match_pairs(db, pairs, 0.6, cap=len(pairs) * F // 4)
        t_put = time.perf_counter() - t0
        # pairs with at least minMatch putative matches go on (computeFeaturesAndMatches.cpp:222-232)
        t0 = time.perf_counter()
        cnt = np.diff(po.astype(np.int64))
        keep = np.flatnonzero(cnt >= 60)
        sel = np.concatenate([np.arange(po[p], po[p + 1]) for p in keep]).astype(np.int64) if len(keep) else np.zeros(0, np.int64)
        pair_of = np.repeat(keep, cnt[keep])
        xI = sc["map_xy"][off[pairs[pair_of, 0]].astype(np.int64) + pi[sel]]
        xJ = sc["map_xy"][off[pairs[pair_of, 1]].astype(np.int64) + pj[sel]]
        goff = np.zeros(len(keep) + 1, np.uint64); goff[1:] = np.cumsum(cnt[keep])
        sizes = np.tile(np.array([w, h, w, h], np.int32), (len(keep), 1))
        t_prep = time.perf_counter() - t0
        g.geometric_filter(xI, xJ, goff, sizes, 4.0, 500, 2)          # warm-up at full size: scratch grows here
        t0 = time.perf_counter()
        r = g.geometric_filter(xI, xJ, goff, sizes, 4.0, 500, 1)
        t_geo = time.perf_counter() - t0
        valid = np.flatnonzero(r["valid"])
        gp = pairs[keep[valid]]
        g.guided_match(db, sc["map_xy"], gp, r["F"][valid], r["error_max"][valid] ** 2, cap=len(gp) * F // 2)   # warm-up
        t0 = time.perf_counter()
        go, gi, gj = g.guided_match(db, sc["map_xy"], gp, r["F"][valid], r["error_max"][valid] ** 2, cap=len(gp) * F // 2)
        t_gm = time.perf_counter() - t0
        t0 = time.perf_counter()
        g.guided_match(db, sc["map_xy"], gp, r["F"][valid], r["error_max"][valid] ** 2, dedup=False, cap=len(gp) * F // 2)
        t_gm_nodedup = time.perf_counter() - t0
        db.free()
    # the same three stages on the CPU port (test oracle, OpenMP) for a few pairs, extrapolated
    from oracle import oracle as orc
    orc.build()
    n_cpu = 6
    tp = tg = tm = 0.0
    for (I, J) in pairs[:n_cpu]:
        a = sc["rows"][int(off[I]):int(off[I + 1])]; b = sc["rows"][int(off[J]):int(off[J + 1])]
        xa = sc["map_xy"][int(off[I]):int(off[I + 1])]; xb = sc["map_xy"][int(off[J]):int(off[J + 1])]
        t0 = time.perf_counter(); oi, oj = orc.match_pair(a, b, 0.6); tp += time.perf_counter() - t0
        if len(oi) < 60:
            continue
        t0 = time.perf_counter(); rr = orc.fmatrix_acransac(xa[oi], xb[oj], (w, h), (w, h), 4.0, 500, 1); tg += time.perf_counter() - t0
        if rr["ok"]:
            t0 = time.perf_counter(); orc.guided_match(rr["F"], xa, a, xb, b, rr["error_max"] ** 2, 0.36); tm += time.perf_counter() - t0
    scale = len(pairs) / n_cpu
    cpu = {"cores": orc.num_threads(), "kind": "port", "sample": "%d pairs, extrapolated to %d" % (n_cpu, len(pairs)),
           "putative_s": tp * scale, "geometric_filter_s": tg * scale, "guided_matching_s": tm * scale,
           "total_s": (tp + tg + tm) * scale}
    print(json.dumps({"cpu_baseline": cpu, "workload": "%d views x %d features, %d pairs" % (V, F, len(pairs)), "putative_s": t_put,
                      "putative_gdist_per_s": len(pairs) * F * F / t_put / 1e9, "putative_matches": int(len(pi)),
                      "pairs_with_60_matches": int(len(keep)), "host_gather_s": t_prep,
                      "geometric_filter_s": t_geo, "pairs_valid": int(len(valid)),
                      "inliers": int(r["n_inliers"].sum()), "guided_matching_s": t_gm, "guided_matching_without_position_dedup_s": t_gm_nodedup, "guided_matches": int(len(gi)),
                      "total_gpu_stages_s": t_put + t_geo + t_gm, "scene_generation_s": gen}), flush=True)


if __name__ == "__main__":
    main()
