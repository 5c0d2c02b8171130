"""Golden vectors for the resection half: P3P solution sets from OpenCV.

Run here (CPU container, cv2 4.13):   python tests/golden/make_golden_resect.py
OpenMVG 1.1 (the reference's resection dependency) is neither vendored nor installed, and the
reference has no test for it, so resection parity is UNPINNED against the reference itself.
What can be pinned is the minimal solver: every correct P3P solver returns the same finite
solution set for a triplet, so cv2.solveP3P (SOLVEPNP_P3P and SOLVEPNP_AP3P, which must agree)
provides the expected poses for 64 seeded triplets."""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from sfmlocalization_b200 import synth  # noqa: E402


def main():
    K = synth.K_IPHONE6
    tri_x, tri_X, sols, nsol, truth = [], [], [], [], []
    for s in range(64):
        sc = synth.resection_scene(3, 500 + s, outlier_frac=0.0, noise_px=0.0)
        x, X = sc["x2d"], sc["X3d"]
        res = []
        for flag in (cv2.SOLVEPNP_P3P, cv2.SOLVEPNP_AP3P):
            n, rvs, tvs = cv2.solveP3P(X, x, K, None, flags=flag)
            res.append(sorted([np.c_[cv2.Rodrigues(rv)[0], tv.reshape(3, 1)] for rv, tv in zip(rvs, tvs)],
                              key=lambda m: tuple(np.round(m.ravel(), 6))))
        a, b = res
        if len(a) != len(b) or any(np.abs(p - q).max() > 1e-6 for p, q in zip(a, b)):
            continue   # solvers disagree on a near-degenerate triplet: not a usable vector
        pad = np.full((4, 3, 4), np.nan)
        for k, m in enumerate(a[:4]):
            pad[k] = m
        tri_x.append(x); tri_X.append(X); sols.append(pad); nsol.append(len(a))
        truth.append(np.c_[sc["R"], sc["t"].reshape(3, 1)])
    out = dict(x2d=np.array(tri_x), X3d=np.array(tri_X), solutions=np.array(sols), n_solutions=np.array(nsol),
               truth=np.array(truth), K=K)
    np.savez_compressed(os.path.join(HERE, "resect_golden.npz"), **out)
    print("wrote resect_golden.npz with", len(tri_x), "triplets; solutions per triplet:", np.bincount(nsol))


if __name__ == "__main__":
    main()
