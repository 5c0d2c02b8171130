"""Batched re-resection of all views of a reconstruction (OpenMVG_BA's first stage,
adjust_sfm_data.cpp:91-146): hulo_resect_acransac_batch against a loop of hulo_resect_acransac.
One JSON line per configuration; wall time of the C-ABI call with host buffers."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from sfmlocalization_b200 import synth  # noqa: E402
from sfmlocalization_b200.gpu import HuloGpu  # noqa: E402


def main():
    ctx = HuloGpu(0)
    cases = [(200, 500, 0.3), (1000, 500, 0.3), (2000, 300, 0.5), (2000, 1500, 0.4), (256, 2000, 0.6)]
    if len(sys.argv) > 1:
        cases = [tuple(float(v) if "." in v else int(v) for v in a.split(",")) for a in sys.argv[1:]]
    for n_views, N, outl in cases:
        scs = [synth.resection_scene(N, 10_000 + i, outlier_frac=outl) for i in range(n_views)]
        off = np.arange(n_views + 1, dtype=np.uint64) * N
        x2d = np.concatenate([s["x2d"] for s in scs]); X3d = np.concatenate([s["X3d"] for s in scs])
        K = scs[0]["K"]
        ctx.resect_acransac_batch(x2d, X3d, off, K, 4096, seed=1)          # scratch growth
        l0 = ctx.launch_count
        t0 = time.perf_counter()
        got = ctx.resect_acransac_batch(x2d, X3d, off, K, 4096, seed=1)
        t_batch = time.perf_counter() - t0
        l_batch = ctx.launch_count - l0
        n_loop = min(n_views, 200)
        t0 = time.perf_counter()
        for p in range(n_loop):
            ctx.resect_acransac(scs[p]["x2d"], scs[p]["X3d"], K, 4096, 1 + 1000003 * p)
        t_loop = (time.perf_counter() - t0) * n_views / n_loop
        print(json.dumps(dict(stage="batched view resection", views=n_views, points_per_view=N, outlier_frac=outl,
                              max_iter=4096, batch_ms=round(t_batch * 1e3, 2),
                              us_per_view=round(t_batch * 1e6 / n_views, 1), launches=l_batch,
                              found=int(sum(g["found"] for g in got)),
                              loop_of_single_calls_ms=round(t_loop * 1e3, 1), loop_sample=n_loop)), flush=True)


if __name__ == "__main__":
    main()
