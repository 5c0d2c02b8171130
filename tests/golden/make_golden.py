"""Generates the committed golden vectors of the matching path with OpenCV's exact matchers.

Run here (CPU container, cv2 4.13):   python tests/golden/make_golden.py
The reference repository ships no fixtures for this path and its own
matcher (OpenCV 3.0 FLANN behind VisionLocalizeCommon/src/MatchUtils.cpp:105-108) cannot be
built here, so the exact 2-NN is pinned on the two OpenCV implementations that define the
"(distance, index) ascending" order the reference's result containers assume:
cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) and cv2.flann_Index(LINEAR, FLANN_DIST_HAMMING).
Both must agree before a vector is written.  The LSH vector records what the reference's
approximate configuration (LshIndexParams(2, 20, 2), checks=2, MatchUtils.cpp:52-65) returns
on the same input: used only for the "-1 / INT_MAX when not found" convention and recall
reporting, never as the parity target.
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from sfmlocalization_b200 import synth  # noqa: E402

FLANN_DIST_HAMMING = 9


def exact_knn2(A, B):
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    m = bf.knnMatch(A, B, k=2)
    idx = np.array([[x.trainIdx for x in r] for r in m], np.int32)
    dist = np.array([[int(x.distance) for x in r] for r in m], np.int32)
    fl = cv2.flann_Index(B, dict(algorithm=0), FLANN_DIST_HAMMING)   # LINEAR
    fi, fd = fl.knnSearch(A, 2, params=dict(checks=-1, sorted=True))
    assert (fi.astype(np.int32) == idx).all() and (fd.astype(np.int32) == dist).all(), "cv2 matchers disagree"
    return idx, dist


def main():
    out = {}
    # 1. tie-heavy: only 12 bits vary, hundreds of d0 == d1 rows
    B = synth.tie_heavy_rows(3000, 11)
    A = synth.tie_heavy_rows(500, 12)
    idx, dist = exact_knn2(A, B)
    assert (dist[:, 0] == dist[:, 1]).sum() > 100
    out.update(tie_A=A, tie_B=B, tie_idx=idx, tie_dist=dist)
    # 2. planted matches in random rows
    A, B, target = synth.descriptor_sets(300, 2000, 21)
    idx, dist = exact_knn2(A, B)
    hit = target >= 0
    assert (idx[hit, 0] == target[hit]).all()
    out.update(pl_A=A, pl_B=B, pl_idx=idx, pl_dist=dist, pl_target=target.astype(np.int32))
    # 3. 61-byte rows as AKAZE emits them (FileUtils.cpp:77-92 pads them to 64)
    A61 = np.ascontiguousarray(synth.random_rows(64, 31)[:, :61])
    B61 = np.ascontiguousarray(synth.random_rows(512, 32)[:, :61])
    idx, dist = exact_knn2(A61, B61)
    out.update(w61_A=A61, w61_B=B61, w61_idx=idx, w61_dist=dist)
    # 4. duplicates: database rows repeated, so d0 == d1 == 0 and the lower index must win
    B = synth.random_rows(200, 41)
    B = np.concatenate([B, B[:50]], axis=0)
    A = B[np.arange(0, 250, 5)].copy()
    idx, dist = exact_knn2(A, B)
    out.update(dup_A=A, dup_B=B, dup_idx=idx, dup_dist=dist)
    # 5. the reference's approximate configuration on vector 2 (convention + recall only)
    A, B = out["pl_A"], out["pl_B"]
    lsh = cv2.flann_Index(B, dict(algorithm=6, table_number=2, key_size=20, multi_probe_level=2),
                          FLANN_DIST_HAMMING)
    li, ld = lsh.knnSearch(A, 2, params=dict(checks=2, eps=0.0, sorted=True))
    out.update(lsh_idx=li.astype(np.int32), lsh_dist=ld.astype(np.int32))
    # 6. ratio-test edge cases evaluated in numpy float32: (0.0f + d0) / d1 < ratio && d1 < INT_MAX
    cases = []
    for ratio in (0.6, 0.7, 0.8):
        r32 = np.float32(ratio)
        for d1 in (1, 2, 5, 10, 50, 100, 243, 486, 512, 2**31 - 1):
            for d0 in sorted({0, 1, int(d1 * ratio) - 1, int(d1 * ratio), int(d1 * ratio) + 1, d1}):
                if d0 < 0 or d0 > d1:
                    continue
                with np.errstate(divide="ignore", invalid="ignore"):
                    q = (np.float32(0.0) + np.float32(d0)) / np.float32(d1)
                cases.append((d0, d1, ratio, int(bool(q < r32) and d1 < 2**31 - 1)))
        cases.append((0, 0, ratio, 0))     # 0/0 -> NaN -> rejected
    out["ratio_cases"] = np.array(cases, np.float64)
    np.savez_compressed(os.path.join(HERE, "matching_golden.npz"), **out)
    print("wrote matching_golden.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
