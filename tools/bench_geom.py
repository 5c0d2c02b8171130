"""K3 timing: hulo_geometric_filter over batches of pairs (wall clock of the C-ABI call incl.
H2D/D2H, and device time from the context's CUDA-event timer).  One JSON line per shape."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from sfmlocalization_b200 import synth  # noqa: E402
from sfmlocalization_b200.gpu import HuloGpu  # noqa: E402


def make(P, N, outl, seed):
    xs, ys, off = [], [], [0]
    for k in range(P):
        tv = synth.two_view_matches(N, seed + k, outlier_frac=outl)
        xs.append(tv["xI"]); ys.append(tv["xJ"]); off.append(off[-1] + N)
    w, h = synth.IMAGE_WH
    return np.concatenate(xs), np.concatenate(ys), np.array(off, np.uint64), np.tile(np.array([w, h, w, h], np.int32), (P, 1))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--one":        # one launch of one shape (ncu capture)
        P, N, outl, rounds = int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5])
        xI, xJ, off, sizes = make(P, N, outl, 1)
        with HuloGpu(0) as g:
            for _ in range(2):
                r = g.geometric_filter(xI, xJ, off, sizes, 4.0, rounds, 1)
        print("valid", int(r["valid"].sum()), "of", P)
        return
    shapes = [(64, 40, 0.5, 25), (64, 200, 0.5, 25), (256, 100, 0.5, 25), (1024, 100, 0.5, 25), (64, 200, 0.5, 200),
              (148, 1000, 0.5, 500), (592, 1000, 0.5, 500), (592, 1000, 1.0, 500), (2000, 300, 0.6, 500)]
    with HuloGpu(0) as g:
        for P, N, outl, rounds in shapes:
            xI, xJ, off, sizes = make(P, N, outl, 1)
            g.geometric_filter(xI, xJ, off, sizes, 4.0, rounds, 1)          # warm-up
            wall, dev = [], []
            for rep in range(5):
                g.timer_start()
                t0 = time.perf_counter()
                r = g.geometric_filter(xI, xJ, off, sizes, 4.0, rounds, 1 + rep)
                wall.append((time.perf_counter() - t0) * 1e3)
                dev.append(g.timer_stop())
            print(json.dumps(dict(kernel="K3 fmatrix_acransac", pairs=P, matches_per_pair=N, outlier_frac=outl,
                                  ransac_round=rounds, wall_ms=round(float(np.median(wall)), 3),
                                  device_ms=round(float(np.median(dev)), 3),
                                  us_per_pair=round(float(np.median(dev)) * 1e3 / P, 2), valid=int(r["valid"].sum()))),
                  flush=True)


if __name__ == "__main__":
    main()
