"""Golden vectors for the F-matrix geometric filter: 7-point solution sets from OpenCV.

Run here (CPU container, cv2 4.13):   python tests/golden/make_golden_fmatrix.py
OpenMVG 1.1 (GeometricFilter_FMatrix_AC, called by hulo::geometricMatch,
VisionLocalizeCommon/src/MatchUtils.cpp:372-420) is neither vendored nor installed and the
reference has no test for it, so parity of the filter is UNPINNED against the reference.  The
minimal solver can be pinned: for seven correspondences every correct 7-point solver returns
the same (1 or 3) fundamental matrices up to scale, so cv2.findFundamentalMat(FM_7POINT)
provides the expected solution sets for 48 seeded samples of preconditioned points."""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from sfmlocalization_b200 import synth  # noqa: E402


def main():
    w, h = synth.IMAGE_WH
    s = 1.0 / np.sqrt(w * h)
    x1s, x2s, sols, nsol = [], [], [], []
    for seed in range(48):
        tv = synth.two_view_matches(7, 900 + seed, outlier_frac=0.0, noise_px=0.3)
        x1 = (tv["xI"] - np.array([w / 2, h / 2])) * s
        x2 = (tv["xJ"] - np.array([w / 2, h / 2])) * s
        F, _ = cv2.findFundamentalMat(x1, x2, cv2.FM_7POINT)
        if F is None:
            continue
        F = F.reshape(-1, 3, 3)
        F = np.array([f / np.linalg.norm(f) for f in F])
        pad = np.full((3, 3, 3), np.nan)
        pad[:len(F)] = F
        x1s.append(x1); x2s.append(x2); sols.append(pad); nsol.append(len(F))
    np.savez_compressed(os.path.join(HERE, "fmatrix_golden.npz"), x1=np.array(x1s), x2=np.array(x2s),
                        solutions=np.array(sols), n_solutions=np.array(nsol))
    print("wrote fmatrix_golden.npz with", len(x1s), "samples; solutions per sample:", np.bincount(nsol))


if __name__ == "__main__":
    main()
