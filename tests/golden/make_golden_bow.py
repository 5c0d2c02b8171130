"""Golden vectors for the BoF view selection (hulo::selectViewByBoF, BoFUtils.cpp:27-68): exact L2
k nearest rows from OpenCV's brute-force matcher, plus what the reference's own FLANN configuration
(KD-tree, 4 trees, 64 checks) returns on the same data, for context (it is approximate).
    python tests/golden/make_golden_bow.py"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(11)
    n, d, knn = 400, 257, 20
    centres = rng.random((12, d)).astype(np.float32)
    bof = (centres[rng.integers(0, 12, n)] + 0.15 * rng.random((n, d))).astype(np.float32)
    bof /= bof.sum(axis=1, keepdims=True)                     # L1-normalised histograms
    queries = (bof[rng.integers(0, n, 6)] + 0.01 * rng.random((6, d))).astype(np.float32)
    bf = cv2.BFMatcher(cv2.NORM_L2)
    exact = np.array([[m.trainIdx for m in ms] for ms in bf.knnMatch(queries, bof, k=knn)], np.int32)
    fl = cv2.FlannBasedMatcher(dict(algorithm=1, trees=4), dict(checks=64))
    fl.add([bof]); fl.train()
    approx = np.array([[m.trainIdx for m in ms] for ms in fl.knnMatch(queries, k=knn)], np.int32)
    np.savez_compressed(os.path.join(HERE, "bow_golden.npz"), bof=bof, queries=queries, exact=exact, flann_kdtree=approx)
    print("recall of the reference's KD-tree configuration:", np.mean([len(set(a) & set(e)) / knn for a, e in zip(approx, exact)]))


if __name__ == "__main__":
    main()
