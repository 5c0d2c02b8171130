"""K1g timing: hulo_guided_match over image pairs of a 200 x 5000 collection (positions uniform in
a 1920 x 1080 image, one synthetic two-view F for every pair, gate 2 px, ratio 0.36).
Prints one JSON line per batch size; wall clock of the C-ABI call (incl. position upload, compaction,
D2H, host de-duplication)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from sfmlocalization_b200 import synth  # noqa: E402
from sfmlocalization_b200.gpu import HuloGpu  # noqa: E402


def main():
    n_img, rows = 200, 5000
    if "--rows" in sys.argv:                      # larger images: fewer of them
        k = sys.argv.index("--rows")
        rows = int(sys.argv[k + 1]); n_img = 24
        del sys.argv[k:k + 2]
    allrows, off = synth.image_collection(n_img, rows, 2000, overlap=0.3)
    rng = np.random.default_rng(5)
    xy = np.stack([rng.uniform(0, 1920, n_img * rows), rng.uniform(0, 1080, n_img * rows)], axis=1)
    F = synth.two_view_matches(8, 1)["F_true"]
    pl = [(a, b) for a in range(n_img) for b in range(a + 1, n_img)]
    counts = [int(a) for a in sys.argv[1:]] or [200, 2000]
    with HuloGpu(0) as g:
        db = g.db(allrows, off)
        for P in counts:
            pairs = pl[:P]
            Fs = np.tile(F, (P, 1, 1)); thr = np.full(P, 4.0)
            g.guided_match(db, xy, pairs, Fs, thr, cap=P * rows)
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                o, gi, gj = g.guided_match(db, xy, pairs, Fs, thr, cap=P * rows)
                ts.append(time.perf_counter() - t0)
            dt = float(np.median(ts))
            print(json.dumps({"kernel": "K1g guided_kernel", "pairs": P, "rows_per_image": rows, "wall_ms": dt * 1e3,
                              "gate_tests_per_s": P * rows * rows / dt, "us_per_pair": dt / P * 1e6,
                              "matches": int(len(gi))}), flush=True)
        db.free()


if __name__ == "__main__":
    main()
