#!/usr/bin/env python
"""Generates tests/golden/anchors_golden.npz: inlier sets computed by OpenCV's own robust estimators
(cv2.solvePnPRansac with the P3P minimal solver, cv2.findFundamentalMat FM_RANSAC) on the seeded
synthetic scenes of sfmlocalization_b200/synth.py, at the inlier threshold the AC-RANSAC restatement
(oracle/oracle_resect.c, oracle_match.c) estimated for the same scene.

Purpose: OpenMVG 1.1 -- where SfM_Localizer::Localize (LocalizeEngine.cc:531) and
GeometricFilter_FMatrix_AC (MatchUtils.cpp:407-416) live -- cannot be obtained here, so the
oracle's AC-RANSAC is anchored on an INDEPENDENT implementation of the same estimation problem:
both must agree on which correspondences are inliers (Jaccard index, tests/test_oracle_anchors.py),
and so must the GPU path (tests/test_resect_gpu.py, test_geom_gpu.py).

Run in the build container (needs cv2):  python tests/golden/make_golden_anchors.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from sfmlocalization_b200 import synth  # noqa: E402

RESECT_CASES = [(100, 0.3, 501), (500, 0.5, 502), (2000, 0.7, 503), (300, 0.1, 504), (1000, 0.6, 505)]
FMAT_CASES = [(300, 0.3, 601), (800, 0.5, 602), (2000, 0.4, 603), (150, 0.2, 604)]


def main():
    orc.build()
    out = {"resect_cases": np.array(RESECT_CASES, np.float64), "fmat_cases": np.array(FMAT_CASES, np.float64)}
    for k, (N, outl, seed) in enumerate(RESECT_CASES):
        sc = synth.resection_scene(N, seed, outlier_frac=outl)
        r = orc.acransac(sc["x2d"], sc["X3d"], sc["K"], max_iter=4096, seed=1)
        assert r["ok"]
        thr = float(r["error_max"])
        cv2.setRNGSeed(seed)
        ok, rvec, tvec, inl = cv2.solvePnPRansac(sc["X3d"], sc["x2d"], sc["K"], None, flags=cv2.SOLVEPNP_P3P,
                                                 reprojectionError=thr, iterationsCount=4096, confidence=0.99999)
        assert ok
        mask = np.zeros(N, bool); mask[inl.ravel()] = True
        mine = np.zeros(N, bool); mine[r["inliers"]] = True
        j = (mask & mine).sum() / (mask | mine).sum()
        print("resection N=%d outliers=%.1f: threshold %.3f px, oracle %d, cv2 %d, truth %d, Jaccard %.4f"
              % (N, outl, thr, mine.sum(), mask.sum(), sc["inlier_mask"].sum(), j))
        out["resect_%d_threshold_px" % k] = np.float64(thr)
        out["resect_%d_cv2_inliers" % k] = mask
    for k, (N, outl, seed) in enumerate(FMAT_CASES):
        sc = synth.two_view_matches(N, seed, outlier_frac=outl)
        r = orc.fmatrix_acransac(sc["xI"], sc["xJ"], sc["size"], sc["size"], 16.0, 1024, 1)
        assert r["ok"]
        thr = float(r["error_max"])
        cv2.setRNGSeed(seed)
        F, m = cv2.findFundamentalMat(sc["xI"], sc["xJ"], cv2.FM_RANSAC, ransacReprojThreshold=thr, confidence=0.99999,
                                      maxIters=4096)
        mask = m.ravel().astype(bool)
        mine = np.zeros(N, bool); mine[r["inliers"]] = True
        j = (mask & mine).sum() / (mask | mine).sum()
        print("F-matrix N=%d outliers=%.1f: threshold %.3f px, oracle %d, cv2 %d, truth %d, Jaccard %.4f"
              % (N, outl, thr, mine.sum(), mask.sum(), sc["inlier_mask"].sum(), j))
        out["fmat_%d_threshold_px" % k] = np.float64(thr)
        out["fmat_%d_cv2_inliers" % k] = mask
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "anchors_golden.npz"), **out)


if __name__ == "__main__":
    main()
