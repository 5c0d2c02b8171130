#!/usr/bin/env python
"""Small invocations of every kernel family through the C-ABI, for compute-sanitizer
(one --tool per gpurun call; see tools/run_sanitizer.sh).  Sizes are chosen to exercise the
multi-chunk / ring wrap-around paths of K1 and K1t while staying fast under the tool."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sfmlocalization_b200 import synth  # noqa: E402
from sfmlocalization_b200.gpu import HuloGpu, LocalizeEngine  # noqa: E402


def main():
    which = sys.argv[1:] or ["k1", "k1t", "items", "k2", "k3", "k1g"]
    with HuloGpu(0) as g:
        A, B, _ = synth.descriptor_sets(700, 40000, 1)
        if "k1" in which:
            g.set_knn_engine("int")
            i0, d0 = g.knn2_host(A, B)                       # several chunks per tile, ring wraps many times
            print("k1 ok", int(d0[:, 0].sum()))
        if "k1t" in which:
            g.set_knn_engine("tc")
            i1, d1 = g.knn2_host(A, B)                       # 6 searcher tiles x chunks, partial last tile
            i2, d2 = g.knn2_host(A[:130], B[:300])
            print("k1t ok", int(d1[:, 0].sum()))
        if "items" in which:
            rows, off = synth.image_collection(6, 600, 2, overlap=0.3)
            query, _ = synth.plant_matches(synth.random_rows(500, 3), rows, 4, frac=0.5)
            db = g.db(rows, off)
            for eng in ("int", "tc"):
                g.set_knn_engine(eng)
                m = g.match_to_query(db, query, 0.6)
                o, oi, oj = g.match_pairs(db, [(0, 1), (1, 2), (2, 5)], 0.7)
                print("items", eng, len(m["i"]), len(oi))
            db.free()
        g.set_knn_engine("auto")
        if "k2" in which:
            rs = synth.resection_scene(300, 41, outlier_frac=0.5)
            r = g.resect_acransac(rs["x2d"], rs["X3d"], rs["K"], max_iter=512, seed=7)
            r2 = g.resect_acransac(rs["x2d"], rs["X3d"], rs["K"], max_iter=512, seed=7, sequential=True)
            print("k2 ok", len(r["inliers"]), len(r2["inliers"]))
        if "k3" in which or "k1g" in which:
            sc = synth.localization_scene(2, 400, 600, 10, 31, track_frac=0.7)
            so = sc["seg_offsets"]
            xys = [sc["map_xy"][int(so[v]):int(so[v + 1])] for v in range(2)]
            n = min(len(xys[0]), len(xys[1]), 300)
            wh = synth.IMAGE_WH
            tv = synth.two_view_matches(300, 5, outlier_frac=0.4)
            r = g.geometric_filter(tv["xI"], tv["xJ"], np.array([0, 300], np.uint64),
                                   np.array([[wh[0], wh[1], wh[0], wh[1]]], np.int32), 4.0, 100, 3)
            print("k3 ok", int(r["n_inliers"][0]))
            if "k1g" in which and r["valid"][0]:
                db = g.db(sc["rows"], so)
                go, gi, gj = g.guided_match(db, sc["map_xy"], [(0, 1)], r["F"][:1], np.array([16.0]))
                db.free()
                print("k1g ok", len(gi))


if __name__ == "__main__":
    main()
