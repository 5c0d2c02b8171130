"""K4 timing: hulo_ransac_transform3d with ransacRound = 100 x #matches (mergeSfM.py:577), against
the numpy restatement of the reference loop (pinned identical to the reference's own function) on a
bounded number of rounds.  One JSON line per shape."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle import oracle as orc  # noqa: E402
from sfmlocalization_b200.gpu import HuloGpu  # noqa: E402
from tests.test_oracle_merge import similarity_case  # noqa: E402


def main():
    with HuloGpu(0) as g:
        for n in (100, 500, 2000):
            A, B, M, inl, _ = similarity_case(20 + n, n, 0.5, 0.01)
            rounds = 100 * n
            for sim in (False, True):
                g.ransac_transform3d(A, B, 0.06, 1000, 1.75, similarity=sim)
                ts = []
                for rep in range(3):
                    t0 = time.perf_counter()
                    Mh, got = g.ransac_transform3d(A, B, 0.06, rounds, 1.75, similarity=sim, seed=rep)
                    ts.append(time.perf_counter() - t0)
                rng = np.random.default_rng(1)
                cpu_rounds = 2000
                samples = np.array([rng.choice(n, 4, replace=False) for _ in range(cpu_rounds)])
                t0 = time.perf_counter()
                orc.ransac_transform3d(A, B, 0.06, samples, 1.75, similarity=sim)
                cpu = (time.perf_counter() - t0) / cpu_rounds * rounds
                print(json.dumps({"kernel": "K4 ransac_transform3d", "flavour": "similarity" if sim else "affine",
                                  "matches": n, "rounds": rounds, "gpu_wall_ms": float(np.median(ts)) * 1e3,
                                  "rounds_per_s": rounds / float(np.median(ts)),
                                  "cpu_numpy_loop_ms_extrapolated": cpu * 1e3, "cpu_sample_rounds": cpu_rounds,
                                  "inliers": int(len(got)), "max_abs_err_M": float(np.abs(Mh - M).max())}), flush=True)


if __name__ == "__main__":
    main()
