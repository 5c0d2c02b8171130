#!/usr/bin/env python
"""bench.py -- benchmarks of the SfMLocalization hot path on B200, one JSON line per run.

Headline (default, --workload c3 = BASELINE.json configs[2], the one the roofline target is quoted
on): exact Hamming 2-NN of 4096 query descriptors against a 10M-row map table of 64-byte AKAZE/MLDB
rows, synthetic random descriptors with planted true matches.  At N > 1 the table is row-sharded
across the ranks (one process per GPU), every rank returns its local top-2 and the candidates
(16 bytes per query per rank) are exchanged and merged (strong scaling: the table is fixed).

Other configurations of BASELINE.json behind --workload, same JSON contract:
    c1  one query localisation, 2000 descriptors vs a 200k-descriptor map (match + AC-RANSAC)
    c2  exhaustive pairwise matching, 200 images x 5000, pair list sharded over the ranks
    c4  batched server, 256 queries x 3000 vs a 2M-descriptor map, queries sharded over the ranks
    c5  campus map, 16384 queries x 50M rows, row-sharded (needs 8 GPUs for the headline shape)

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cX]
                    [--engine tc|int]

No PyTorch: the library owns device memory, its stream, CUDA events and the communicator;
torchrun is only the process launcher (RANK / LOCAL_RANK / WORLD_SIZE).  Every CPU leg (the
reference arm's work, cpu_baseline, the parity sample) runs in a child process, so the process that
drives the GPU never maps the oracle.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from sfmlocalization_b200 import synth  # noqa: E402

METRIC = "hamming_2nn_gdist_per_s"
UNIT = "Gdist/s"
N_QUERIES = 4096
N_MAP = 10_000_000
SEED = 3000          # seed = 1000 * config number
BLK = 250_000        # map rows are generated in independently seeded blocks
PARITY_QUERIES = 64  # rows of the result checked bit for bit against the oracle over the whole table


def read_json(path, default=None):
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return default


def popc_peak_gdist(sm_mhz_max):
    """Integer-popcount roofline: SMs x POPC lanes/clk/SM x f / 16 POPC per distance.
    Uses the lanes/clk measured by tools/microbench on this pool's B200 when committed under
    profiles/popc_peak.json, else the nominal 16 lanes/clk/SM."""
    m = read_json(os.path.join(ROOT, "profiles", "popc_peak.json"), {}) or {}
    lanes = float(m.get("popc_lanes_per_clk_per_sm", 16.0))
    sms = int(m.get("sms", 148))
    src = "measured (tools/microbench, profiles/popc_peak.json)" if "popc_lanes_per_clk_per_sm" in m \
        else "nominal 16 POPC lanes/clk/SM"
    return sms * lanes * sm_mhz_max * 1e6 / 16.0 / 1e9, src


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.mark = 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def mark_timed_region(self):
        """Samples before this point belong to the warm-up."""
        self.mark = len(self.lines)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        lines = self.lines[self.mark:] if len(self.lines) - self.mark >= 3 else self.lines
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for k, name in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "power_w_max": float(max(power)) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def rendezvous_id(rank, world, make_id):
    """Share rank 0's 128-byte NCCL id through a file keyed by the launcher (single node)."""
    key = "%s_%s" % (os.environ.get("MASTER_PORT", "0"), os.getppid())
    path = "/tmp/hulo_nccl_id_%s" % key
    if rank == 0:
        uid = make_id()
        tmp = path + ".tmp"
        with open(tmp, "wb") as f:
            f.write(uid)
        os.replace(tmp, path)
        return uid, path
    t0 = time.time()
    while time.time() - t0 < 300:
        if os.path.exists(path) and os.path.getsize(path) == 128:
            with open(path, "rb") as f:
                return f.read(), path
        time.sleep(0.05)
    raise RuntimeError("rank %d: no NCCL id at %s" % (rank, path))


def map_block(b):
    return synth.random_rows(BLK, SEED + 1 + b)


def make_queries():
    """The query rows: identical on every rank.  30 % are noisy copies of rows of map block 0."""
    A = synth.random_rows(N_QUERIES, SEED + 7919)
    return synth.plant_matches(A, map_block(0), SEED + 104729, frac=0.3)


def make_tables(world, rank):
    """Every rank generates the same query set; rank r generates only its shard of the map
    (blocks are seeded independently, so shards concatenate to the 1-GPU table)."""
    n_blocks = N_MAP // BLK
    if n_blocks % world != 0:
        raise SystemExit("--gpus must divide %d" % n_blocks)
    per = n_blocks // world
    b_lo, b_hi = rank * per, (rank + 1) * per
    shard = np.concatenate([map_block(b) for b in range(b_lo, b_hi)], axis=0)
    A, target = make_queries()
    return A, target, shard, b_lo * BLK


# --------------------------------------------------------------------------- CPU legs (child process)
def run_child(kind, extra=(), timeout=900):
    """Runs `bench.py --cpu-leg kind` in a child process with the launcher's OMP_NUM_THREADS removed;
    returns (its JSON line, path of the npz it wrote or None)."""
    fd, out = tempfile.mkstemp(prefix="hulo_cpu_leg_", suffix=".npz")
    os.close(fd)
    env = dict(os.environ)
    for k in ("OMP_NUM_THREADS", "OMP_PROC_BIND", "OMP_PLACES", "GOMP_CPU_AFFINITY"):
        env.pop(k, None)
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-leg", kind, "--out", out,
           "--workload", ARGS.workload] + list(extra)
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("cpu leg %s failed: %s" % (kind, r.stderr[-2000:]))
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    return json.loads(line), out


def oracle_all_cores():
    from oracle import oracle as orc
    orc.build()
    return orc, orc.use_all_cores()


def calibrate_queries(orc, A, shard, grp, target_s):
    """Number of query rows (a multiple of the port's 8-rows-per-thread group) for which one
    exact pass over `shard` takes about target_s on this host: doubling probes until a probe
    runs >= 0.5 s (short probes under-read the rate while the OpenMP pool and clocks ramp up)."""
    nq = grp
    while True:
        t0 = time.perf_counter()
        orc.knn2(A[:nq], shard)
        dt = time.perf_counter() - t0
        if dt >= 0.5 or nq >= A.shape[0]:
            break
        nq = min(A.shape[0], nq * 2)
    want = nq * target_s / max(dt, 1e-6)
    return int(min(A.shape[0], max(grp, want // grp * grp)))


def leg_knn2(args):
    """cpu_baseline of the flat search: the oracle port (exact 2-NN, every host core) on a bounded
    sample: the first 1M map rows, as many queries as ~12 s allow.  Writes the sample's result."""
    orc, threads = oracle_all_cores()
    A, _ = make_queries()
    shard = np.concatenate([map_block(b) for b in range(4)], axis=0)
    grp = 8 * max(1, threads)                       # the port walks 8 searcher rows per thread
    nq = calibrate_queries(orc, A, shard, grp, args.budget)
    nb = int(shard.shape[0])
    t0 = time.perf_counter()
    idx, dist = orc.knn2(A[:nq], shard)
    dt = time.perf_counter() - t0
    np.savez(args.out, idx=idx, dist=dist, nq=nq, nb=nb)
    print(json.dumps({"value": nq * nb / dt / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
                      "sample": "%d queries x %d map rows (%.1f s), exact 2-NN, oracle/oracle_match.c with OpenMP"
                                % (nq, nb, dt)}))


def leg_parity(args):
    """Exact top-2 of the first PARITY_QUERIES queries over the WHOLE map table (all blocks, whatever
    the sharding), block by block with the (distance, global index) merge."""
    orc, threads = oracle_all_cores()
    A, _ = make_queries()
    A = A[:PARITY_QUERIES]
    cd, ci = [], []
    for b in range(N_MAP // BLK):
        i, d = orc.knn2(A, map_block(b))
        ci.append(i.astype(np.int64) + b * BLK); cd.append(d.astype(np.int64))
    cd = np.concatenate(cd, axis=1); ci = np.concatenate(ci, axis=1)
    order = np.lexsort((ci, cd), axis=1)[:, :2]
    np.savez(args.out, idx=np.take_along_axis(ci, order, axis=1).astype(np.int32),
             dist=np.take_along_axis(cd, order, axis=1).astype(np.int32))
    print(json.dumps({"queries": PARITY_QUERIES, "map_rows": N_MAP, "cores": threads}))


def matching_dtype(g):
    """Arithmetic type of the 2-NN behind the matchers on this context: the tensor-core engine's
    operand type when the engine setting lets searches of this size use it, else the integer pipes."""
    if g.knn_engine == "int":
        return "u32"
    if g.knn_engine == "tc8":
        return "s8"
    return "fp4" if os.environ.get("HULO_TC_BITS", "4") != "8" else "s8"


def c1_scene():
    return synth.localization_scene(100, 2000, 20000, 2000, 1000)


def leg_localize(args):
    """CPU port of one C1 query: per-view exact 2-NN + ratio on all cores, assembly, sequential
    AC-RANSAC on one thread (as the reference runs it); the geometric filter timed on its own."""
    orc, threads = oracle_all_cores()
    sc = c1_scene()
    t0 = time.perf_counter()
    off = sc["seg_offsets"]
    m_view, m_i, m_j, m_d = [], [], [], []
    for v in range(len(off) - 1):
        oi, oj, od = orc.match_view_to_query(sc["rows"][int(off[v]):int(off[v + 1])], sc["q_desc"], 0.6)
        if len(oi) < 16:
            continue
        m_view += [v] * len(oi); m_i += oi.tolist(); m_j += oj.tolist(); m_d += od.tolist()
    t1 = time.perf_counter()
    w, h = synth.IMAGE_WH
    mv, mi, mj = np.array(m_view), np.array(m_i), np.array(m_j)
    n_geo_valid = 0
    for p, v in enumerate(sorted(set(m_view))):
        sel = np.nonzero(mv == v)[0]
        rr = orc.fmatrix_acransac(sc["map_xy"][int(off[v]) + mi[sel]], sc["q_xy"][mj[sel]], (w, h), (w, h), 4.0, 25,
                                  77 + 1000003 * p)
        n_geo_valid += int(rr["ok"])
    t1g = time.perf_counter()
    order = np.lexsort((sc["obs_feat"], sc["obs_view"]))
    cj, cl = orc.match_set(m_view, m_i, m_j, m_view, m_j, m_d, sc["obs_view"][order], sc["obs_feat"][order],
                           sc["obs_landmark"][order].astype(np.int64), len(sc["q_desc"]))
    t2 = time.perf_counter()
    ro = orc.acransac(sc["q_xy"][cj], sc["landmark_X"][cl], sc["K"], max_iter=4096, seed=1)
    t3 = time.perf_counter()
    out = {"ms_per_query": ((t1 - t0) + (t3 - t1g)) * 1e3, "cores": threads, "kind": "port",
           "stage_ms": {"putMatch": (t1 - t0) * 1e3, "assembly": (t2 - t1g) * 1e3, "PnP": (t3 - t2) * 1e3,
                        "geoMatch_when_enabled": (t1g - t1) * 1e3},
           "sample": "1 query: exact per-view 2-NN on all cores, sequential AC-RANSAC on one thread (as the "
                     "reference runs it); geoMatch = F-matrix AC-RANSAC of the kept views one after the other on "
                     "one thread, 25 rounds",
           "localized": bool(ro["ok"]), "inliers": int(len(ro["inliers"])), "geometric_pairs_valid": n_geo_valid,
           "reference_algorithm_lsh": lsh_reference(sc, set(zip(m_view, m_i, m_j)))}
    print(json.dumps(out))


def c2_collection():
    return synth.image_collection(200, 5000, 2000, overlap=0.3)


def leg_pairs(args):
    """CPU port of C2 on a bounded sample: the first pairs of the all-pairs list, orc_match_pair
    (exact 2-NN of image I in image J, ratio 0.7, one-to-one filter), one pair after the other,
    the 2-NN of each pair on all cores."""
    orc, threads = oracle_all_cores()
    rows, off = c2_collection()
    pairs = [(a, b) for a in range(200) for b in range(a + 1, 200)]
    n, t0, done, nm = 0, time.perf_counter(), 0.0, 0
    out_i = []
    while done < args.budget and n < len(pairs):
        a, b = pairs[n]
        i, j = orc.match_pair(rows[int(off[a]):int(off[a + 1])], rows[int(off[b]):int(off[b + 1])], 0.7)
        out_i.append(np.stack([np.full(len(i), n), i, j], axis=1))
        nm += len(i); n += 1
        done = time.perf_counter() - t0
    np.savez(args.out, matches=np.concatenate(out_i, axis=0).astype(np.int64), n_pairs=n)
    print(json.dumps({"value": n * 25e6 / done / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
                      "pairs_per_s": n / done,
                      "sample": "first %d of 19900 pairs (%.1f s), 5000 x 5000 exact 2-NN + ratio 0.7 + one-to-one "
                                "filter per pair (oracle orc_match_pair), %d matches" % (n, done, nm)}))


def c4_scene():
    return synth.localization_scene(1000, 2000, 200000, 3000, 4100, window=6000)


def leg_server(args):
    """CPU port of C4 on a bounded sample: whole queries of the batched-server workload, one after
    the other: per-view exact 2-NN + ratio over the 1000 views on all cores, assembly, AC-RANSAC."""
    orc, threads = oracle_all_cores()
    sc = c4_scene()
    off = sc["seg_offsets"]
    order = np.lexsort((sc["obs_feat"], sc["obs_view"]))
    n, t0, done, ok = 0, time.perf_counter(), 0.0, 0
    while done < args.budget and n < 256:
        q = synth.extra_query(sc, 3000, 4100 + 10 + n)
        m_view, m_i, m_j, m_d = [], [], [], []
        for v in range(len(off) - 1):
            oi, oj, od = orc.match_view_to_query(sc["rows"][int(off[v]):int(off[v + 1])], q["q_desc"], 0.6)
            if len(oi) < 16:
                continue
            m_view += [v] * len(oi); m_i += oi.tolist(); m_j += oj.tolist(); m_d += od.tolist()
        cj, cl = orc.match_set(m_view, m_i, m_j, m_view, m_j, m_d, sc["obs_view"][order], sc["obs_feat"][order],
                               sc["obs_landmark"][order].astype(np.int64), len(q["q_desc"]))
        ro = orc.acransac(q["q_xy"][cj], sc["landmark_X"][cl], sc["K"], max_iter=4096, seed=1) if len(cj) > 8 \
            else {"ok": False}
        ok += int(bool(ro["ok"])); n += 1
        done = time.perf_counter() - t0
    print(json.dumps({"value": n / done, "unit": "localizations/s", "cores": threads, "kind": "port",
                      "sample": "%d of 256 queries (%.1f s): 3000 descriptors vs 2M map rows in 1000 views, exact "
                                "2-NN on all cores, sequential AC-RANSAC on one thread; %d localised" % (n, done, ok)}))


def lsh_reference(sc, exact_matches):
    """The reference's own (approximate) matcher configuration, for context: cv::flann LSH index
    (2 tables, key 20, multi-probe 2) on the query descriptors, knnSearch(k=2, checks=2) per map
    view, float ratio test (MatchUtils.cpp:52-65, 303-355), through OpenCV's Python binding,
    one thread (the reference spreads the views over OpenMP threads).  Reports its time and how
    its putative matches compare with the exact matcher's.  Not the parity target."""
    try:
        import cv2
    except Exception as e:                      # pragma: no cover
        return {"unavailable": "cv2 not importable: %s" % e}
    off = sc["seg_offsets"]
    t0 = time.perf_counter()
    index = cv2.flann_Index(sc["q_desc"], dict(algorithm=6, table_number=2, key_size=20, multi_probe_level=2), 9)
    got = set()
    for v in range(len(off) - 1):
        a = sc["rows"][int(off[v]):int(off[v + 1])]
        idx, dist = index.knnSearch(a, 2, params=dict(checks=2, eps=0.0, sorted=True))
        with np.errstate(divide="ignore", invalid="ignore"):
            q = dist[:, 0].astype(np.float32) / dist[:, 1].astype(np.float32)
        keep = np.nonzero((q < np.float32(0.6)) & (dist[:, 1] < 2**31 - 1))[0]
        if len(keep) >= 16:
            got.update((v, int(i), int(idx[i, 0])) for i in keep)
    dt = time.perf_counter() - t0
    both = len(got & exact_matches)
    return {"ms_per_query_matching": dt * 1e3, "threads": 1, "putative_matches": len(got),
            "exact_matcher_matches": len(exact_matches), "common": both,
            "note": "approximate and indicative only: on i.i.d. synthetic descriptors the probed LSH buckets "
                    "almost never hold a second candidate, so rows come back with d1 = INT_MAX and are rejected "
                    "(MatchUtils.cpp:349); on real images the second candidate is whatever shares a bucket"}


CPU_LEGS = {"knn2": leg_knn2, "parity": leg_parity, "localize": leg_localize, "pairs": leg_pairs,
            "server": leg_server}


# --------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path for this stage.  Its own matcher (OpenCV 3.0 FLANN
    behind MatchUtils.cpp:105-108) and resection (OpenMVG 1.1) cannot be built in this image, so the
    oracle port is timed, with every host thread (the launcher's OMP_NUM_THREADS is ignored), on
    bounded samples of the same workload.  Rank 0 alone works."""
    if rank != 0:
        return
    base = {"impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "vs_baseline": None, "data": "synthetic", "gpu_launches": 0}
    if args.workload in ("c3", "c5"):
        orc, threads = oracle_all_cores()
        A, _ = make_queries()
        shard = np.concatenate([map_block(b) for b in range(4)], axis=0)
        grp = 8 * max(1, threads)
        nb = int(shard.shape[0])
        nq = calibrate_queries(orc, A, shard, grp, 2.0)    # one step ~ 2 s of CPU work
        for _ in range(args.warmup):
            orc.knn2(A[:nq], shard)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            orc.knn2(A[:nq], shard)
        dt = time.perf_counter() - t0
        value = nq * nb * args.steps / dt / 1e9
        sample = "%d queries x %d map rows per step, exact 2-NN, oracle port (OpenMP)" % (nq, nb)
        base.update({"metric": METRIC, "value": value, "unit": UNIT, "ms_per_step": dt / args.steps * 1e3,
                     "scaling": "strong", "dtype": "u32", "config": workload_config(args.gpus, "cpu", "none"),
                     "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                     "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(base))
        return
    # c1 / c2 / c4: one bounded CPU sample is the whole run (steps and warm-up do not multiply it)
    kind = {"c1": "localize", "c2": "pairs", "c4": "server"}[args.workload]
    cb, out = run_child(kind, ["--budget", "20"])
    if os.path.exists(out):
        os.remove(out)
    if args.workload == "c1":
        value, unit, metric = 1e3 / cb["ms_per_query"], "localizations/s", "query_localizations_per_s"
        cb = {"value": value, "unit": unit, "cores": cb["cores"], "kind": "port", "sample": cb["sample"],
              "stage_ms": cb["stage_ms"]}
    elif args.workload == "c2":
        value, unit, metric = cb["value"], UNIT, METRIC
    else:
        value, unit, metric = cb["value"], "localizations/s", "query_localizations_per_s"
    base.update({"metric": metric, "value": value, "unit": unit, "ms_per_step": None, "scaling": "weak",
                 "dtype": "u32 / f64", "config": {"workload": WORKLOAD_NAMES[args.workload]},
                 "cpu_baseline": cb,
                 "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base))


# --------------------------------------------------------------------------- C1: one query
def localize_bench(g, with_cpu=True, reps=20):
    """Second half of the metric (BASELINE.json: query localizations/sec): configs[0], one query
    image of 2000 descriptors against a 200k-descriptor map (100 views x 2000), ratio 0.6
    (the server's secondTestRatio, localizeImage.cc:46-59), AC-RANSAC resection with the
    reference's 4096-iteration budget.  End to end through hulo_engine_localize: query
    descriptors and keypoints in host memory, pose back in host memory."""
    from sfmlocalization_b200.gpu import LocalizeEngine
    sc = c1_scene()
    eng = LocalizeEngine(g, sc["rows"], sc["seg_offsets"], sc["obs_view"], sc["obs_feat"], sc["obs_landmark"],
                         sc["landmark_X"], sc["K"], ratio=0.6)
    for k in range(3):
        r = eng.localize(sc["q_desc"], sc["q_xy"], seed=k)
    wall, stages, ok, err = [], [], 0, []
    launches0 = g.launch_count
    for k in range(reps):
        t0 = time.perf_counter()
        r = eng.localize(sc["q_desc"], sc["q_xy"], seed=100 + k)
        wall.append((time.perf_counter() - t0) * 1e3)
        stages.append(r["times_ms"])
        if r["localized"]:
            ok += 1
            err.append(float(np.linalg.norm(r["center"] - sc["center"])))
    launches_per_query = (g.launch_count - launches0) / max(reps, 1)
    # the same with the F-matrix geometric filter between matching and assembly
    # (hulo::geometricMatch, LocalizeEngine.cc:458; ransacRound 25, precision 4 px: LocalizeParam.py:35)
    eng.set_keypoints(sc["map_xy"], sc["view_wh"], synth.IMAGE_WH)
    eng.configure_geometric(True, 25, 4.0)
    gwall, gstages, gok, gerr = [], [], 0, []
    eng.localize(sc["q_desc"], sc["q_xy"], seed=7)
    for k in range(reps):
        t0 = time.perf_counter()
        rg = eng.localize(sc["q_desc"], sc["q_xy"], seed=200 + k)
        gwall.append((time.perf_counter() - t0) * 1e3)
        gstages.append(rg["times_ms"])
        if rg["localized"]:
            gok += 1
            gerr.append(float(np.linalg.norm(rg["center"] - sc["center"])))
    # throughput form of the same shape: 64 different query images of the same camera in one
    # hulo_engine_localize_batch call (one matching pass over the map, one batched resection), no filter
    eng.configure_geometric(False)
    qs = [synth.extra_query(sc, 2000, 500 + k) for k in range(64)]
    qd = [q["q_desc"] for q in qs]; qx = [q["q_xy"] for q in qs]
    eng.localize_batch(qd[:4], qx[:4], seed=1)
    eng.localize_batch(qd, qx, seed=2)
    t0 = time.perf_counter()
    bt = eng.localize_batch(qd, qx, seed=3)
    b_dt = time.perf_counter() - t0
    b_err = [float(np.linalg.norm(bt["center"][k] - qs[k]["center"])) for k in range(64) if bt["localized"][k]]
    eng.close()
    gst = np.median(np.array(gstages), axis=0)
    ms = float(np.median(wall))
    st = np.median(np.array(stages), axis=0)
    out = {"workload": WORKLOAD_NAMES["c1"],
           "ms_per_query": ms, "localizations_per_s": 1e3 / ms,
           "stage_ms": {"putMatch": float(st[0]), "assembly": float(st[1]), "PnP": float(st[2])},
           "fraction_localized": ok / reps, "centre_error_m_median": float(np.median(err)) if err else None,
           "correspondences": int(len(r["corr_qfeat"])), "inliers": int(len(r["inliers"])),
           "kernel_launches_per_query": launches_per_query,
           "target_ms": 5.0,
           "with_geometric_filter": {
               "ms_per_query": float(np.median(gwall)), "localizations_per_s": 1e3 / float(np.median(gwall)),
               "stage_ms": {"putMatch": float(gst[0]), "geoMatch": float(gst[3]), "assembly": float(gst[1]),
                            "PnP": float(gst[2])},
               "settings": "F-matrix AC-RANSAC per (view, query) pair, ransacRound 25, precision 4 px",
               "fraction_localized": gok / reps,
               "centre_error_m_median": float(np.median(gerr)) if gerr else None,
               "correspondences": int(len(rg["corr_qfeat"])), "inliers": int(len(rg["inliers"]))},
           "batched_64_queries": {
               "ms_per_query": b_dt * 1e3 / 64, "localizations_per_s": 64 / b_dt,
               "stage_ms_total": {"putMatch": float(bt["times_ms"][0]), "assembly": float(bt["times_ms"][1]),
                                  "PnP": float(bt["times_ms"][2])},
               "fraction_localized": float(bt["localized"].mean()),
               "centre_error_m_median": float(np.median(b_err)) if b_err else None,
               "note": "64 query images of 2000 descriptors in one hulo_engine_localize_batch call, host buffers in, poses out"}}
    if with_cpu:
        cb, tmp = run_child("localize")
        if os.path.exists(tmp):
            os.remove(tmp)
        out["reference_algorithm_lsh"] = cb.pop("reference_algorithm_lsh", None)
        out["cpu_baseline"] = cb
    return out


WORKLOAD_NAMES = {
    "c1": "C1: 2000 query descriptors vs 200000 map descriptors (100 views), ratio 0.6, AC-RANSAC + P3P, max 4096 "
          "iterations",
    "c2": "C2: exhaustive pairwise matching, 200 images x 5000 descriptors, 19900 pairs, ratio 0.7, one-to-one filter",
    "c4": "C4 batched server: 256 queries x 3000 descriptors vs 2M-descriptor map (1000 views), matching + AC-RANSAC "
          "resection (4096-iteration budget), end to end",
}


def workload_config(n_gpus, engine, exchange):
    name = "C5 campus-scale map" if N_MAP > 10_000_000 else "C3 building-scale map"
    four = os.environ.get("HULO_TC_BITS", "4") != "8"
    img = " + %.1f GB %s tile image" % (N_MAP * (256 if four else 512) / 1e9, "fp4" if four else "int8") if engine == "tc" else ""
    return {"workload": "%s: %d queries x %d map descriptors (64-byte AKAZE/MLDB rows), exact Hamming 2-NN, planted "
                        "matches (30%%)" % (name, N_QUERIES, N_MAP),
            "engine": {"tc": ("K1t4: fp4 (e2m1 +-1) contraction on tcgen05 tensor cores, kind::mxf4 with unit block scales, "
                              "fp32 accumulation (distance = (512 - dot) / 2, exact)") if four else
                             "K1t: int8 contraction on tcgen05 tensor cores (distance = (512 - dot) / 2, int32 exact)",
                       "int": "K1: XOR + popcount on the integer pipes", "cpu": "oracle port on host cores"}[engine],
            "sharding": "single GPU, whole table resident" if n_gpus == 1 else
                        "map rows sharded over %d GPUs; local top-2 per rank, exchange = %s, merge on every rank"
                        % (n_gpus, {"peer-store": "stores into peer-mapped buffers over NVLink (CUDA IPC) fused "
                                                  "into the chunk-merge kernel + flag wait; NCCL only for set-up",
                                    "nccl-allgather": "one ncclAllGather of 16 bytes per query per rank",
                                    "none": "none"}.get(exchange, exchange)),
            "cache": "map table %d MB%s > 126 MB L2, streamed every step (no L2 flush needed)" % (N_MAP * 64 // 10**6, img)}


# --------------------------------------------------------------------------- C3 / C5: flat search
def bench_flat(args, rank, world, local_rank):
    from sfmlocalization_b200.gpu import HuloGpu, PinnedArray
    import ctypes as C

    A, target, shard, row_base = make_tables(world, rank)
    g = HuloGpu(local_rank)
    id_path = None
    if world > 1:
        uid, id_path = rendezvous_id(rank, world, HuloGpu.comm_unique_id)
        g.comm_init(uid, rank, world)
    dA = g.db(A)
    dB = g.db(shard)
    nA = A.shape[0]
    total_dist = float(nA) * float(N_MAP)

    def step_device():
        if world > 1:
            g.knn2_sharded(dA, dB, row_base, fetch=False)
        else:
            g.knn2(dA, dB, fetch=False)

    def timed_device(engine, steps, warmup, sampler=None):
        g.set_knn_engine(engine)
        for _ in range(max(warmup, 3)):
            step_device()
        g.comm_barrier() if world > 1 else g.synchronize()
        if sampler:
            sampler.mark_timed_region()
        l0 = g.launch_count
        g.timer_start()
        for _ in range(steps):
            step_device()
        ms = g.timer_stop()
        launches = g.launch_count - l0
        ms = g.comm_max(ms) if world > 1 else ms
        return ms, launches

    # ---- kernel-resident throughput: inputs already in HBM, results left in HBM.  The clock sampler
    # runs from the warm-up on (nvidia-smi needs ~100 ms for its first line); the samples reported are
    # those of the timed region when it is long enough to hold three.
    primary = args.engine
    other = "int" if primary == "tc" else "tc"
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, launches = timed_device(primary, args.steps, args.warmup, sampler)
    clocks = sampler.stop()
    value = total_dist * args.steps / (ms * 1e-3) / 1e9
    exchange = g.exchange_kind
    # ---- end to end through the C-ABI with host buffers: every step uploads the queries from
    # pinned host memory and reads the top-2 back; the map table is engine state (resident).  Timed
    # right after the device-resident loop, before the other engines heat the board
    pin_A = PinnedArray(A.shape, np.uint8); pin_A.array[...] = A
    pin_i = PinnedArray((nA, 2), np.int32); pin_d = PinnedArray((nA, 2), np.int32)
    lib = g.lib

    def step_e2e():
        rc = lib.hulo_db_update(g.h, dA.h, pin_A.array.ctypes.data_as(C.c_void_p), nA, 64)
        assert rc == 0
        if world > 1:
            rc = lib.hulo_knn2_sharded(g.h, dA.h, dB.h, row_base, pin_i.array.ctypes.data_as(C.c_void_p),
                                       pin_d.array.ctypes.data_as(C.c_void_p))
        else:
            rc = lib.hulo_knn2(g.h, dA.h, dB.h, pin_i.array.ctypes.data_as(C.c_void_p),
                               pin_d.array.ctypes.data_as(C.c_void_p))
        assert rc == 0

    # the same as a pipeline (hulo_knn2_sharded_submit / _collect, the form a server answering batch
    # after batch uses): every step still uploads its queries and reads its own results back inside
    # the timed region, but the results of step k are collected after step k + 1 has been issued, so
    # the exchange and the result copy overlap the next search
    def upload():
        rc = lib.hulo_db_update(g.h, dA.h, pin_A.array.ctypes.data_as(C.c_void_p), nA, 64)
        assert rc == 0

    def submit():
        rc = lib.hulo_knn2_sharded_submit(g.h, dA.h, dB.h, row_base)
        assert rc == 0, lib.hulo_last_error()

    def collect():
        n = C.c_size_t(0)
        rc = lib.hulo_knn2_sharded_collect(g.h, pin_i.array.ctypes.data_as(C.c_void_p),
                                           pin_d.array.ctypes.data_as(C.c_void_p), C.byref(n))
        assert rc == 0 and n.value == nA, lib.hulo_last_error()

    def run_pipelined(steps):
        upload(); submit()
        for _ in range(steps - 1):
            upload(); submit(); collect()
        collect()

    step_e2e()
    g.comm_barrier() if world > 1 else g.synchronize()
    e2e_sampler = ClockSampler(local_rank)               # over both end-to-end loops
    e2e_sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    g.synchronize()
    serial_s = time.perf_counter() - t0
    serial_s = g.comm_max(serial_s) if world > 1 else serial_s
    serial_idx, serial_dist = pin_i.array.copy(), pin_d.array.copy()
    run_pipelined(2)
    g.comm_barrier() if world > 1 else g.synchronize()
    t0 = time.perf_counter()
    run_pipelined(args.steps)
    g.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_clocks = e2e_sampler.stop()
    e2e_s = g.comm_max(e2e_s) if world > 1 else e2e_s
    e2e_pipelined_value = total_dist * args.steps / e2e_s / 1e9
    e2e_serial_value = total_dist * args.steps / serial_s / 1e9
    # the headline e2e is the call pattern a host would choose: the pipeline where there is an
    # exchange to hide (N > 1), the plain blocking call on one GPU (nothing to overlap but a 64 KB copy)
    e2e_value = e2e_pipelined_value if world > 1 else e2e_serial_value
    idx, dist = pin_i.array.copy(), pin_d.array.copy()
    pipeline_equals_serial = bool(np.array_equal(idx, serial_idx) and np.array_equal(dist, serial_dist))

    o_steps = max(3, min(args.steps, 10))
    o_ms, _ = timed_device(other, o_steps, 3)
    engines = {primary: {"value": value, "ms_per_step": ms / args.steps},
               other: {"value": total_dist * o_steps / (o_ms * 1e-3) / 1e9, "ms_per_step": o_ms / o_steps}}
    # the int8 form of the tensor-core engine (K1t), for the record
    t8_ms, _ = timed_device("tc8", o_steps, 3)
    engines["tc8"] = {"value": total_dist * o_steps / (t8_ms * 1e-3) / 1e9, "ms_per_step": t8_ms / o_steps}
    g.set_knn_engine(primary)

    # ---- cold variant: map shard uploaded from host inside the timed call (hulo_knn2_host)
    cold = None
    if world == 1 and args.workload == "c3":
        t0 = time.perf_counter()
        g.knn2_host(A, shard)
        cold_s = time.perf_counter() - t0
        cold = {"value": total_dist / cold_s / 1e9, "unit": UNIT,
                "note": "one call of hulo_knn2_host: 640 MB map + queries H2D from pageable memory inside the call"}

    # ---- checks on the result of the timed configuration
    ok = True
    hit = np.nonzero(target >= 0)[0]
    ok &= bool(np.array_equal(idx[hit, 0], target[hit]))
    ok &= bool((dist[:, 0] <= dist[:, 1]).all())
    # both engines must return the same arrays
    engines_agree = True
    for eng in (other, "tc8"):
        g.set_knn_engine(eng)
        step_device()
        oi, od = g.knn2_fetch(nA)
        engines_agree &= bool(np.array_equal(oi, idx) and np.array_equal(od, dist))
    g.set_knn_engine(primary)
    ok &= engines_agree
    ok &= pipeline_equals_serial

    parity = None
    if rank == 0:
        # bit-equality of a fixed query subset with the oracle over the WHOLE table, at every N
        pj, ppath = run_child("parity")
        want = np.load(ppath)
        parity = bool(np.array_equal(idx[:PARITY_QUERIES], want["idx"]) and
                      np.array_equal(dist[:PARITY_QUERIES], want["dist"]))
        os.remove(ppath)
        ok &= parity

    if rank == 0:
        peaks = read_json(os.path.join(ROOT, "MEASURED_PEAKS.json"), {}) or {}
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        sm_max = clocks.get("sm_max_mhz") or 1965.0
        per_gpu = value / world
        step_s = ms * 1e-3 / args.steps
        int_peak, int_src = popc_peak_gdist(sm_max)
        # the limit of K1's own instruction mix: per distance 26.5 LOP3 + 3 VIMNMX
        # on the ALU pipe (2 cycles per warp instruction per sub-partition) and 7.75 POPC on the XU pipe
        # (8 cycles each); the busier pipe bounds the rate
        mix_cycles = max((26.5 + 3.0) * 2.0, 7.75 * 8.0)
        mix_peak = 148 * 4 * 32 * sm_max * 1e6 / mix_cycles / 1e9
        int_gpu = engines["int"]["value"] / world
        int_roof = {"bound": "int-popc", "achieved": int_gpu, "peak": int_peak, "unit": UNIT + "/GPU",
                    "frac": int_gpu / int_peak, "frac_of_binding_pipe": int_gpu / mix_peak,
                    "binding_pipe_peak": mix_peak, "peak_source": int_src,
                    "mix": "26.5 LOP3 + 3 VIMNMX (ALU, 2 clk) | 7.75 POPC (XU, 8 clk) | 7.75 IMAD (FMA)",
                    "work_per_unit": "1 dist = 512 compared bits = 16 x 32-bit POPC (naive); the kernel folds words "
                                     "with LOP3 carry-save adders first, so frac can exceed 1"}
        tc_gpu = engines["tc"]["value"] / world
        bf16_s = float(peaks.get("bf16_tflops_sustained", 1400.0)); bf16_b = float(peaks.get("bf16_tflops", 1590.0))
        tc_bits = 8 if os.environ.get("HULO_TC_BITS", "4") == "8" else 4
        tc_mult = 4.0 if tc_bits == 4 else 2.0      # nominal dense rate of the operand type over bf16
        tc_tops = tc_gpu * 1024.0 / 1e3             # 1 dist = 512 multiply-adds = 1024 ops
        tprof = read_json(os.path.join(ROOT, "profiles", "k1t4_traffic.json" if tc_bits == 4 else "k1t_traffic.json"), {}) or {}
        tc_roof = {"bound": "tensor", "achieved": tc_tops, "peak": tc_mult * bf16_s, "unit": "TOP/s",
                   "frac": tc_tops / (tc_mult * bf16_s),
                   "peak_source": "%g x bf16_tflops_sustained of MEASURED_PEAKS.json: no %s GEMM was measured on this "
                                  "pool, the nominal dense rate of the operand type is %g x the bf16 rate; sustained "
                                  "because the launches run back to back under the 1 kW cap"
                                  % (tc_mult, "fp4" if tc_bits == 4 else "int8", tc_mult) if peaks else "fallback",
                   "frac_of_burst_peak": tc_tops / (tc_mult * bf16_b),
                   "frac_of_nominal": tc_tops / (9000.0 if tc_bits == 4 else 4500.0),
                   "tensor_pipe_active_pct_ncu": tprof.get("sm__pipe_tensor_cycles_active_pct"),
                   "operands": "e2m1 (fp4) values +-1, ue8m0 block scales 1.0, fp32 accumulation (kind::mxf4)" if tc_bits == 4
                               else "int8 values +-1, int32 accumulation (kind::i8)",
                   "work_per_unit": "1 dist = 512 multiply-adds = 1024 ops; the accumulated integers stay below 2^10, so "
                                    "the result is exact in either accumulator type"}
        if primary == "tc":
            alg_bytes = (256.0 if tc_bits == 4 else 512.0) * (nA + N_MAP / world) + 16.0 * nA
            roofline = dict(tc_roof)
            roofline["traffic"] = tprof.get("dram_bytes_per_launch")
            roofline["int_engine"] = int_roof
        else:
            alg_bytes = 64.0 * (nA + N_MAP / world) + 16.0 * nA
            roofline = dict(int_roof)
            roofline["traffic"] = (read_json(os.path.join(ROOT, "profiles", "k1_traffic.json"), {}) or {}).get(
                "dram_bytes_per_launch")
            roofline["tc_engine"] = tc_roof
        hbm_gbs = alg_bytes / step_s / 1e9
        roofline["hbm"] = {"achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                           "algorithmic_bytes_per_launch": alg_bytes,
                           "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None,
            "dtype": ("fp4" if os.environ.get("HULO_TC_BITS", "4") != "8" else "s8") if primary == "tc" else "u32",
            "data": "synthetic",
            "config": workload_config(world, primary, exchange),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(nA * 64),
                    "d2h_bytes_per_step": int(nA * 16), "clocks": e2e_clocks,
                    "inputs": "queries H2D from pinned memory + top-2 D2H every step; map table resident "
                              "(uploaded once at engine construction, as LocalizeEngine loads its map once)",
                    "api": ("hulo_db_update + hulo_knn2_sharded_submit / hulo_knn2_sharded_collect: the results of "
                            "step k are collected after step k + 1 has been issued") if world > 1 else
                           "hulo_db_update + hulo_knn2 (blocking), every step",
                    "pipelined_value": e2e_pipelined_value,
                    "blocking_calls_value": e2e_serial_value,
                    "pipeline_equals_blocking_calls": pipeline_equals_serial,
                    "note": "wall clock around the C-ABI calls.  With the tensor-core engine the board sits at its "
                            "1 kW cap, and the copies and synchronisations between steps are idle time the next "
                            "launch gets back as clock, so this rate can exceed the back-to-back device-resident one"},
            "e2e_cold": cold,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "engines": engines,
            "engines_agree": engines_agree,
            "engines_note": "tc = K1t4 (fp4 operands, kind::mxf4), tc8 = K1t (int8 operands, kind::i8), int = K1 (integer pipes, "
                            "the configuration BASELINE.json:north_star prescribes; --engine int makes it the headline); "
                            "identical result arrays",
            "parity_vs_oracle_sample": parity,
            "result_check": "planted matches found, d0<=d1, engines agree, %d queries bit-equal to the oracle over "
                            "all %d rows" % (PARITY_QUERIES, N_MAP) if ok else "FAILED",
        }
        if world == 1 and args.workload == "c3":
            line["localize"] = localize_bench(g, with_cpu=not args.no_cpu_baseline)
        if not args.no_cpu_baseline and world == 1:
            cb, cpath = run_child("knn2", ["--budget", "12"])
            s = np.load(cpath)
            nq, nb = int(s["nq"]), int(s["nb"])
            gi, gd = g.knn2_host(A[:nq], shard[:nb])      # the same sample through the GPU path
            cb["gpu_equals_cpu_on_sample"] = bool(np.array_equal(gi, s["idx"]) and np.array_equal(gd, s["dist"]))
            os.remove(cpath)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    dA.free(); dB.free()
    pin_A.free(); pin_i.free(); pin_d.free()
    if world > 1:
        g.comm_barrier()
    g.close()
    if rank == 0 and id_path and os.path.exists(id_path):
        os.remove(id_path)
    if not ok:
        sys.exit(1)


# --------------------------------------------------------------------------- C1 on its own
def bench_c1(args, rank, world, local_rank):
    """One GPU: one query at a time through hulo_engine_localize.  N > 1: the same query with the VIEWS
    sharded over the ranks (hulo_engine_localize_sharded: every rank matches its range of views, one
    all-gather of the surviving matches, every rank finishes the query); a latency play, so the value
    is the rate of ONE query stream, and rank 0 checks that the sharded result is bit-identical to its
    own single-GPU result."""
    from sfmlocalization_b200.gpu import HuloGpu, LocalizeEngine
    g = HuloGpu(local_rank)
    id_path = None
    if world > 1:
        uid, id_path = rendezvous_id(rank, world, HuloGpu.comm_unique_id)
        g.comm_init(uid, rank, world)
    sampler = ClockSampler(local_rank); sampler.start()
    if world == 1:
        loc = localize_bench(g, with_cpu=not args.no_cpu_baseline, reps=max(args.steps, 10))
        value, ms, steps, extra = loc["localizations_per_s"], loc["ms_per_query"], max(args.steps, 10), {}
    else:
        sc = c1_scene()
        eng = LocalizeEngine(g, sc["rows"], sc["seg_offsets"], sc["obs_view"], sc["obs_feat"], sc["obs_landmark"],
                             sc["landmark_X"], sc["K"], ratio=0.6)
        for k in range(5):
            eng.localize_sharded(sc["q_desc"], sc["q_xy"], seed=k)
        steps = max(args.steps, 20)
        wall, stages = [], []
        launches_q = 0
        for k in range(steps):
            g.comm_barrier()
            l0 = g.launch_count
            t0 = time.perf_counter()
            r = eng.localize_sharded(sc["q_desc"], sc["q_xy"], seed=100 + k)
            dt = (time.perf_counter() - t0) * 1e3
            launches_q = g.launch_count - l0
            wall.append(g.comm_max(dt))
            stages.append(r["times_ms"])
        single = eng.localize(sc["q_desc"], sc["q_xy"], seed=100 + steps - 1)        # the same seed on one GPU
        same = bool(np.array_equal(single["corr_qfeat"], r["corr_qfeat"]) and
                    np.array_equal(single["corr_landmark"], r["corr_landmark"]) and
                    np.array_equal(single["inliers"], r["inliers"]) and np.array_equal(single["R"], r["R"]) and
                    np.array_equal(single["center"], r["center"]))
        same = bool(-g.comm_max(-float(same)) > 0.5)                                  # on every rank
        eng.close()
        ms = float(np.median(wall))
        st = np.median(np.array(stages), axis=0)
        value = 1e3 / ms
        loc = {"workload": WORKLOAD_NAMES["c1"], "ms_per_query": ms, "localizations_per_s": value,
               "stage_ms": {"putMatch_incl_exchange": float(st[0]), "assembly": float(st[1]), "PnP": float(st[2])},
               "localized": bool(r["localized"]), "centre_error_m": float(np.linalg.norm(r["center"] - sc["center"]))}
        loc["kernel_launches_per_query"] = launches_q
        extra = {"sharded_equals_single_gpu": same}
    clocks = sampler.stop()
    exchange_kind = g.exchange_kind
    if world > 1:
        g.comm_barrier()
    g.close()
    if rank == 0 and id_path and os.path.exists(id_path):
        os.remove(id_path)
    if rank == 0:
        line = {"metric": "query_localizations_per_s", "value": value, "unit": "localizations/s", "n_gpus": world,
                "steps": steps, "warmup": 3 if world == 1 else 5, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "fp4 / f32",
                "data": "synthetic",
                "config": {"workload": WORKLOAD_NAMES["c1"],
                           "sharding": "single GPU" if world == 1 else
                                       "views of the map sharded over %d ranks for ONE query; the surviving matches (12 bytes "
                                       "each) are exchanged by %s, resection on every rank"
                                       % (world, "stores from the compaction's output into every rank's peer-mapped buffer "
                                                 "(NVLink) + flags, one D2H copy per query"
                                          if exchange_kind == "peer-store" else
                                          "one ncclAllGather staged through host buffers")},
                "e2e": {"value": value, "unit": "localizations/s", "h2d_bytes_per_step": 2000 * 64 + 2000 * 16,
                        "d2h_bytes_per_step": 96,
                        "inputs": "query descriptors and keypoints from host memory, pose back, every query"},
                "gpu_launches": int(round(loc["kernel_launches_per_query"] * steps)), "clocks": clocks, "localize": loc, "cpu_baseline": loc.get("cpu_baseline")}
        line.update(extra)
        print(json.dumps(line))
        if extra and not extra["sharded_equals_single_gpu"]:
            sys.exit(1)


# --------------------------------------------------------------------------- C2: pair-sharded matching
def bench_pairs(args, rank, world, local_rank):
    """hulo_match_pairs on this rank's share of the 19900 pairs (LPT partition, no data-path
    collective: every rank holds all descriptors).  A step = the whole pair list once."""
    from sfmlocalization_b200.gpu import HuloGpu
    from tests import hostlib
    g = HuloGpu(local_rank)
    id_path = None
    if world > 1:
        uid, id_path = rendezvous_id(rank, world, HuloGpu.comm_unique_id)
        g.comm_init(uid, rank, world)
    n_img, rows = 200, 5000
    allrows, off = c2_collection()
    pairs = np.array([(a, b) for a in range(n_img) for b in range(a + 1, n_img)], np.uint64)
    # one rank keeps the list order (the CPU sample below is its prefix); more ranks: LPT partition
    pl = pairs if world == 1 else pairs[hostlib.partition_pairs(pairs, np.full(n_img, rows), rank, world)]
    db = g.db(allrows, off)                                     # resident: descriptors uploaded once
    cap = 4 << 20
    sampler = ClockSampler(local_rank); sampler.start()
    for _ in range(max(1, min(args.warmup, 2))):
        g.match_pairs(db, pl, 0.7, cap=cap)
    g.comm_barrier() if world > 1 else g.synchronize()
    sampler.mark_timed_region()
    steps = max(1, min(args.steps, 5))
    l0 = g.launch_count
    t0 = time.perf_counter()
    for _ in range(steps):
        o, oi, oj = g.match_pairs(db, pl, 0.7, cap=cap)
    dt = time.perf_counter() - t0
    launches = g.launch_count - l0
    dt = g.comm_max(dt) if world > 1 else dt
    clocks = sampler.stop()
    dist_total = float(len(pairs)) * rows * rows
    value = dist_total * steps / dt / 1e9
    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
                "warmup": max(1, min(args.warmup, 2)), "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": matching_dtype(g), "data": "synthetic",
                "config": {"workload": WORKLOAD_NAMES["c2"],
                           "sharding": "pair list partitioned over %d rank(s) by n_I x n_J (LPT), descriptors (64 MB) "
                                       "replicated, no data-path collective" % world,
                           "cache": "64 MB of descriptors stay in L2; the pair outputs are compacted on the device"},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(len(pl) * 32),
                        "d2h_bytes_per_step": int(len(oi) * 8 + len(o) * 8),
                        "inputs": "the timed call is the C-ABI call itself: pair list H2D, ratio + one-to-one filters "
                                  "and compaction on the device, match lists D2H; value is therefore the end-to-end "
                                  "number"},
                "gpu_launches": int(launches), "clocks": clocks,
                "pairs_per_s": len(pairs) * steps / dt, "matches_on_rank0": int(len(oi)),
                "roofline": {"bound": "int-popc", "achieved": value / world, "peak": popc_peak_gdist(1965.0)[0],
                             "unit": UNIT + "/GPU", "frac": value / world / popc_peak_gdist(1965.0)[0],
                             "note": "item mode of K1 (integer pipes): per pair 5000 x 5000; wall clock of the whole "
                                     "call, not kernel time"}}
        if not args.no_cpu_baseline and world == 1:
            cb, cpath = run_child("pairs", ["--budget", "15"])
            s = np.load(cpath)
            npairs = int(s["n_pairs"])
            k_of = o[:npairs + 1].astype(np.int64)
            got = np.stack([np.repeat(np.arange(npairs), np.diff(k_of)), oi[:k_of[-1]].astype(np.int64),
                            oj[:k_of[-1]].astype(np.int64)], axis=1)
            cb["gpu_equals_cpu_on_sample"] = bool(np.array_equal(got.astype(np.int64), s["matches"]))
            os.remove(cpath)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    db.free()
    if world > 1:
        g.comm_barrier()
    g.close()
    if rank == 0 and id_path and os.path.exists(id_path):
        os.remove(id_path)


# --------------------------------------------------------------------------- C4: batched server
def bench_server(args, rank, world, local_rank):
    """256 concurrent queries end to end (hulo_engine_localize_batch); queries sharded over the
    ranks, map replicated -- replicas, no data-path collective.  A step = all 256 queries once."""
    from sfmlocalization_b200.gpu import HuloGpu, LocalizeEngine
    g = HuloGpu(local_rank)
    id_path = None
    if world > 1:
        uid, id_path = rendezvous_id(rank, world, HuloGpu.comm_unique_id)
        g.comm_init(uid, rank, world)
    n_queries, nq = 256, 3000
    sc = c4_scene()
    mine = list(range(rank, n_queries, world))
    qs = [synth.extra_query(sc, nq, 4100 + 10 + k) for k in mine]
    eng = LocalizeEngine(g, sc["rows"], sc["seg_offsets"], sc["obs_view"], sc["obs_feat"], sc["obs_landmark"],
                         sc["landmark_X"], sc["K"], ratio=0.6)
    descs = [q["q_desc"] for q in qs]; xys = [q["q_xy"] for q in qs]
    sampler = ClockSampler(local_rank); sampler.start()
    eng.localize_batch(descs[:2], xys[:2])
    eng.localize_batch(descs, xys, seed=4)
    g.comm_barrier() if world > 1 else g.synchronize()
    sampler.mark_timed_region()
    steps = max(1, min(args.steps, 5))
    l0 = g.launch_count
    t0 = time.perf_counter()
    for s in range(steps):
        b = eng.localize_batch(descs, xys, seed=5 + s)
    dt = time.perf_counter() - t0
    launches = g.launch_count - l0
    dt = g.comm_max(dt) if world > 1 else dt
    clocks = sampler.stop()
    n_ok = float(b["localized"].sum())
    err = [float(np.linalg.norm(b["center"][k] - qs[k]["center"])) for k in range(len(qs)) if b["localized"][k]]
    n_ok = -g.comm_max(-n_ok) if world > 1 else n_ok
    value = n_queries * steps / dt
    if rank == 0:
        map_rows = int(sc["rows"].shape[0])
        line = {"metric": "query_localizations_per_s", "value": value, "unit": "localizations/s", "n_gpus": world,
                "steps": steps, "warmup": 2, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": matching_dtype(g) + " / f32", "data": "synthetic",
                "config": {"workload": WORKLOAD_NAMES["c4"],
                           "sharding": "queries over %d rank(s), map (%d rows, %d MB) replicated, no data-path "
                                       "collective" % (world, map_rows, map_rows * 64 // 10**6),
                           "cache": "map 128 MB ~ L2 size; searched once per batch"},
                "e2e": {"value": value, "unit": "localizations/s",
                        "h2d_bytes_per_step": int(len(mine) * nq * (64 + 16)), "d2h_bytes_per_step": int(len(mine) * 104),
                        "inputs": "the timed call is hulo_engine_localize_batch itself: query descriptors and "
                                  "keypoints from host memory, poses back; value is the end-to-end number"},
                "gpu_launches": int(launches), "clocks": clocks,
                "matching_gdist_per_s": float(n_queries) * nq * map_rows * steps / dt / 1e9,
                "stage_ms_last_step": {"putMatch": float(b["times_ms"][0]), "assembly": float(b["times_ms"][1]),
                                       "PnP": float(b["times_ms"][2])},
                "min_localized_on_a_rank": int(n_ok), "queries_per_rank": len(mine),
                "centre_error_m_median_rank0": float(np.median(err)) if err else None,
                "roofline": {"bound": "int-popc", "achieved": float(n_queries) * nq * map_rows * steps / dt / 1e9 / world,
                             "peak": popc_peak_gdist(1965.0)[0], "unit": UNIT + "/GPU",
                             "frac": float(n_queries) * nq * map_rows * steps / dt / 1e9 / world / popc_peak_gdist(1965.0)[0],
                             "note": "whole-call wall clock attributed to the matching work (the dominant stage)"}}
        if not args.no_cpu_baseline and world == 1:
            cb, cpath = run_child("server", ["--budget", "25"], timeout=1800)
            if os.path.exists(cpath):
                os.remove(cpath)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    eng.close()
    if world > 1:
        g.comm_barrier()
    g.close()
    if rank == 0 and id_path and os.path.exists(id_path):
        os.remove(id_path)


ARGS = None


def main():
    global ARGS, N_QUERIES, N_MAP, SEED
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c3", choices=["c1", "c2", "c3", "c4", "c5"],
                    help="c3 (default, the headline): 4096 x 10M flat search; see the module docstring")
    ap.add_argument("--engine", default="tc", choices=["tc", "int"],
                    help="arithmetic of the flat search the headline value is measured on (the other one is "
                         "measured beside it): tc = int8 tensor-core contraction (K1t), int = integer pipes (K1)")
    ap.add_argument("--cpu-leg", default=None, choices=sorted(CPU_LEGS), help=argparse.SUPPRESS)
    ap.add_argument("--out", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--budget", type=float, default=12.0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    ARGS = args
    if args.workload == "c5":
        N_QUERIES, N_MAP, SEED = 16384, 50_000_000, 5000
    if args.cpu_leg:
        CPU_LEGS[args.cpu_leg](args)
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload in ("c3", "c5"):
        bench_flat(args, rank, world, local_rank)
    elif args.workload == "c1":
        bench_c1(args, rank, world, local_rank)
    elif args.workload == "c2":
        bench_pairs(args, rank, world, local_rank)
    else:
        bench_server(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
