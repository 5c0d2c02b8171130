import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np
from sfmlocalization_b200 import synth
from sfmlocalization_b200.gpu import HuloGpu
g = HuloGpu(0)
nmap = 2_000_000
rows = synth.random_rows(nmap, 1)
off = np.arange(0, nmap + 1, 2000, dtype=np.uint64)
db = g.db(rows, off)
for nQ in (4, 24):
    qs = synth.random_rows(3000 * nQ, 2)
    qoff = np.arange(0, 3000 * nQ + 1, 3000, dtype=np.uint64)
    g.match_to_queries(db, qs, qoff, 0.6)
    t0 = time.perf_counter(); g.timer_start()
    r = g.match_to_queries(db, qs, qoff, 0.6)
    ms = g.timer_stop(); wall = (time.perf_counter() - t0) * 1e3
    d = nmap * 3000.0 * nQ
    print("batch nQ", nQ, "device ms", round(ms, 1), "wall ms", round(wall, 1), "Gdist/s dev", round(d / ms / 1e6, 1), "matches", len(r["i"]))
# same work, flat
dq = g.db(synth.random_rows(3000 * 24, 2))
g.knn2(db, dq, fetch=False); g.synchronize(); g.timer_start(); g.knn2(db, dq, fetch=False); ms = g.timer_stop()
print("flat 2M x 72000: ms", round(ms, 1), "Gdist/s", round(nmap * 72000.0 / ms / 1e6, 1))
# single-query path
q1 = synth.random_rows(3000, 3)
g.match_to_query(db, q1, 0.6); g.timer_start(); g.match_to_query(db, q1, 0.6); ms = g.timer_stop()
print("single query 2M x 3000: ms", round(ms, 2), "Gdist/s", round(nmap * 3000.0 / ms / 1e6, 1))
