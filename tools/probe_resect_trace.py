"""Host-time trace of one C1 query's resection (HULO_RESECT_TRACE): where the PnP stage's time goes."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sfmlocalization_b200.gpu import HuloGpu, LocalizeEngine  # noqa: E402

g = HuloGpu(0)
sc = bench.c1_scene()
eng = LocalizeEngine(g, sc["rows"], sc["seg_offsets"], sc["obs_view"], sc["obs_feat"], sc["obs_landmark"],
                     sc["landmark_X"], sc["K"], ratio=0.6)
for k in range(20):
    r = eng.localize(sc["q_desc"], sc["q_xy"], seed=k)
print("correspondences", len(r["corr_qfeat"]), "inliers", len(r["inliers"]), "times_ms", r["times_ms"], flush=True)
sys.stderr.flush()
st = []
for k in range(200):
    r = eng.localize(sc["q_desc"], sc["q_xy"], seed=100 + k)
    st.append(r["times_ms"])
print("median stage ms", np.median(np.array(st), axis=0), flush=True)
eng.close(); g.close()
