import sys, time
import numpy as np
sys.path.insert(0, ".")
from sfmlocalization_b200 import synth
from sfmlocalization_b200.gpu import HuloGpu
with HuloGpu(0) as g:
    for N, o in [(300, 0.5), (1500, 0.4), (700, 0.02), (2000, 0.6)]:
        scs = [synth.resection_scene(N, 10_000 + i, outlier_frac=o) for i in range(20)]
        for sc in scs[:3]:
            g.resect_acransac(sc["x2d"], sc["X3d"], sc["K"], seed=1)
        ts = []
        for p, sc in enumerate(scs):
            l0 = g.launch_count
            t0 = time.perf_counter()
            r = g.resect_acransac(sc["x2d"], sc["X3d"], sc["K"], seed=1 + 1000003 * p)
            ts.append(((time.perf_counter() - t0) * 1e3, g.launch_count - l0))
        print(N, o, "median ms %.3f" % np.median([t for t, _ in ts]), "max %.3f" % max(t for t, _ in ts),
              "launches", sorted(set(l for _, l in ts)))
