"""K1 variant sweep on the bench workload (C3, 4096 x 10M, 1 GPU): one line per variant.
usage: python tools/sweep_k1.py "512,4,8,1" "512,4,89,1" ..."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for spec in sys.argv[1:]:
    t, q, c, o = spec.split(",")
    env = dict(os.environ, HULO_KNN_THREADS=t, HULO_KNN_QPT=q, HULO_KNN_CSA=c, HULO_KNN_OPT=o)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3",
                        "--no-cpu-baseline"], env=env, capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print("threads=%s qpt=%s csa=%s opt=%s  %.1f Gdist/s  %.2f ms/step  %s  sm %s MHz" %
              (t, q, c, o, d["value"], d["ms_per_step"], d["result_check"], d["clocks"]["sm_mhz"]), flush=True)
    except Exception as e:
        print("threads=%s qpt=%s csa=%s opt=%s  FAILED: %s %s" % (t, q, c, o, e, r.stderr[-300:]), flush=True)
