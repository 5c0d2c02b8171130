"""Golden vectors for the 3D-3D model-merge RANSAC FROM THE REFERENCE'S OWN
CODE: ransacAffineTransform of PyVisionLocalizeCommon/src/hulo_sfm/mergeSfM.py:344-388 is plain
numpy, so its source is read from /root/reference, its one Python-2 print statement is blanked, and
it is executed here (CPU container) on seeded inputs; random.sample is wrapped so the 4-point
samples of every round are recorded with the result.

    python tests/golden/make_golden_merge.py        # needs /root/reference; writes merge_golden.npz
"""
import os
import random
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/PyVisionLocalizeCommon/src/hulo_sfm/mergeSfM.py"


def load_reference_function():
    src = open(REF).read()
    a = src.index("def ransacAffineTransform(")
    b = src.index("#\n# TODO : create AC-RANSAC version")
    fn = re.sub(r'print "[^"]*" \+ str\(\w+\)', "pass", src[a:b])
    log = []

    class Rand:                      # the module-level name `random` inside the reference function
        @staticmethod
        def sample(population, k):
            s = random.sample(list(population), k)
            log.append(s)
            return s
    ns = {"np": np, "random": Rand, "sys": sys}
    exec(fn, ns)
    return ns["ransacAffineTransform"], log


def case(seed, n, outlier_frac, noise, scale=1.3):
    rng = np.random.default_rng(seed)
    B = rng.normal(size=(3, n)) * 5
    Q = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    M = np.hstack([scale * Q * np.array([1.0, 1.05, 0.97]), rng.normal(size=(3, 1)) * 3])
    A = M @ np.vstack([B, np.ones((1, n))]) + rng.normal(size=(3, n)) * noise
    out = rng.random(n) < outlier_frac
    A[:, out] = rng.normal(size=(3, int(out.sum()))) * 5
    return A, B


def main():
    fn, log = load_reference_function()
    out = {}
    for k, (seed, n, of, noise, thres, rounds, ratio) in enumerate([
            (1, 200, 0.4, 0.01, 0.05, 1500, 1.75), (2, 60, 0.6, 0.02, 0.1, 3000, 1.75),
            (3, 500, 0.2, 0.005, 0.03, 800, 1.2), (4, 40, 0.97, 0.01, 0.02, 500, 1.75)]):
        A, B = case(seed, n, of, noise)
        del log[:]
        random.seed(100 + k)
        M, inl = fn(A, B, thres, rounds, ratio)
        out["A%d" % k] = A; out["B%d" % k] = B
        out["par%d" % k] = np.array([thres, ratio])
        out["samples%d" % k] = np.array(log, np.int64)
        out["M%d" % k] = np.asarray(M, np.float64)
        out["inl%d" % k] = np.asarray(inl, np.int64)
        print("case", k, "n", n, "rounds", len(log), "inliers", len(inl))
    np.savez_compressed(os.path.join(HERE, "merge_golden.npz"), **out)


if __name__ == "__main__":
    main()
