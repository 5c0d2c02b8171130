"""One C1 query localisation, a few times: the launch list of the per-query path (run under
ncu --metrics gpu__time_duration.sum) and host-side wall times of its stages."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from sfmlocalization_b200 import synth  # noqa: E402
from sfmlocalization_b200.gpu import HuloGpu, LocalizeEngine  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
with HuloGpu(0) as g:
    sc = synth.localization_scene(100, 2000, 20000, 2000, 1000)
    eng = LocalizeEngine(g, sc["rows"], sc["seg_offsets"], sc["obs_view"], sc["obs_feat"], sc["obs_landmark"],
                         sc["landmark_X"], sc["K"], ratio=0.6)
    for k in range(3):
        eng.localize(sc["q_desc"], sc["q_xy"], seed=k)
    wall, st = [], []
    for k in range(reps):
        t0 = time.perf_counter()
        r = eng.localize(sc["q_desc"], sc["q_xy"], seed=100 + k)
        wall.append((time.perf_counter() - t0) * 1e3)
        st.append(r["times_ms"])
    print(json.dumps(dict(wall_ms_median=float(np.median(wall)), stages_ms_median=np.median(np.array(st), axis=0).tolist(),
                          launches_per_query=None)))
    eng.set_keypoints(sc["map_xy"], sc["view_wh"], synth.IMAGE_WH)
    eng.configure_geometric(True, 25, 4.0)
    for guided in (False, True):
        eng.set_guided_matching(guided)
        eng.localize(sc["q_desc"], sc["q_xy"], seed=1)
        wall, st = [], []
        for k in range(reps):
            t0 = time.perf_counter()
            r = eng.localize(sc["q_desc"], sc["q_xy"], seed=200 + k)
            wall.append((time.perf_counter() - t0) * 1e3)
            st.append(r["times_ms"])
        print(json.dumps(dict(geometric_filter=True, guided=guided, wall_ms_median=float(np.median(wall)),
                              stages_ms_median=np.median(np.array(st), axis=0).tolist(), localized=bool(r["localized"]))))
    eng.close()
