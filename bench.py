#!/usr/bin/env python
"""bench.py -- headline benchmark of the SfMLocalization hot path on B200.

Workload (BASELINE.json configs[2], the one the roofline target is quoted on): exact Hamming
2-NN of 4096 query descriptors against a 10M-row map table of 64-byte AKAZE/MLDB rows,
synthetic random descriptors with planted true matches.  At N > 1 the table is row-sharded
across the ranks (one process per GPU), every rank returns its local top-2 and one NCCL
all-gather of 16 bytes per query per rank merges them (strong scaling: the table is fixed).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

No PyTorch: the library owns device memory, its stream, CUDA events and the NCCL
communicator; torchrun is only the process launcher (RANK / LOCAL_RANK / WORLD_SIZE).
"""
import argparse
import json
import os
import subprocess
import sys
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

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except ValueError:
                continue
            for k, name in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
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


def make_tables(world, rank):
    """Every rank generates the same query set; rank r generates only its shard of the map
    (blocks are seeded independently, so shards concatenate to the 1-GPU table)."""
    blk = 250_000                       # rows in 250k-row blocks, seeded per block
    n_blocks = N_MAP // blk
    per = n_blocks // world if n_blocks % world == 0 else None
    if per is None:
        raise SystemExit("--gpus must divide %d" % n_blocks)
    b_lo, b_hi = rank * per, (rank + 1) * per
    shard = np.concatenate([synth.random_rows(blk, SEED + 1 + b) for b in range(b_lo, b_hi)], axis=0)
    row_base = b_lo * blk
    # queries: 30 % are noisy copies of map rows.  The planted sources are drawn from block 0 and
    # written by its owner; the query rows themselves are identical on every rank.
    block0 = synth.random_rows(blk, SEED + 1)
    A = synth.random_rows(N_QUERIES, SEED + 7919)
    A, target = synth.plant_matches(A, block0, SEED + 104729, frac=0.3)
    return A, target, shard, row_base


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


def cpu_baseline_sample(A, shard, budget_s=12.0):
    """Times the oracle port (exact 2-NN, all host cores) on a bounded sample of the workload."""
    from oracle import oracle as orc
    orc.build()
    threads = orc.num_threads()
    grp = 8 * max(1, threads)                       # the port walks 8 searcher rows per thread
    nb = int(shard.shape[0])
    nq = calibrate_queries(orc, A, shard, grp, budget_s)
    t0 = time.perf_counter()
    idx, dist = orc.knn2(A[:nq], shard[:nb])
    dt = time.perf_counter() - t0
    return {"value": nq * nb / dt / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d queries x %d map rows (%.1f s), exact 2-NN, oracle/oracle_match.c with OpenMP"
                      % (nq, nb, dt)}, (idx, dist, nq, nb)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path for this stage.  Its own matcher (OpenCV 3.0
    FLANN behind MatchUtils.cpp:105-108) cannot be built in this image, so the oracle port is
    timed, with every host thread, on bounded samples of the same workload."""
    if rank != 0:
        return
    blk = 250_000
    A = synth.random_rows(N_QUERIES, SEED + 7919)
    shard = np.concatenate([synth.random_rows(blk, SEED + 1 + b) for b in range(4)], axis=0)
    from oracle import oracle as orc
    orc.build()
    threads = orc.num_threads()
    grp = 8 * max(1, threads)                       # the port walks 8 searcher rows per thread
    # size one step to ~2 s of CPU work: all 1M staged map rows, as many queries as that allows
    nb = int(shard.shape[0])
    nq = calibrate_queries(orc, A, shard, grp, 2.0)
    for _ in range(args.warmup):
        orc.knn2(A[:nq], shard[:nb])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.knn2(A[:nq], shard[:nb])
    dt = time.perf_counter() - t0
    value = nq * nb * args.steps / dt / 1e9
    sample = "%d queries x %d map rows per step, exact 2-NN, oracle port (OpenMP)" % (nq, nb)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def localize_bench(g, with_cpu=True, reps=20):
    """Second half of the metric (BASELINE.json: query localizations/sec): configs[0], one query
    image of 2000 descriptors against a 200k-descriptor map (100 views x 2000), ratio 0.6
    (the server's secondTestRatio, localizeImage.cc:46-59), AC-RANSAC resection with the
    reference's 4096-iteration budget.  End to end through hulo_engine_localize: query
    descriptors and keypoints in host memory, pose back in host memory."""
    from sfmlocalization_b200.gpu import LocalizeEngine
    sc = synth.localization_scene(100, 2000, 20000, 2000, 1000)
    eng = LocalizeEngine(g, sc["rows"], sc["seg_offsets"], sc["obs_view"], sc["obs_feat"], sc["obs_landmark"],
                         sc["landmark_X"], sc["K"], ratio=0.6)
    for k in range(3):
        r = eng.localize(sc["q_desc"], sc["q_xy"], seed=k)
    wall, stages, ok, err = [], [], 0, []
    for k in range(reps):
        t0 = time.perf_counter()
        r = eng.localize(sc["q_desc"], sc["q_xy"], seed=100 + k)
        wall.append((time.perf_counter() - t0) * 1e3)
        stages.append(r["times_ms"])
        if r["localized"]:
            ok += 1
            err.append(float(np.linalg.norm(r["center"] - sc["center"])))
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
    out = {"workload": "C1: 2000 query descriptors vs 200000 map descriptors (100 views), ratio 0.6, "
                       "AC-RANSAC + P3P, max 4096 iterations",
           "ms_per_query": ms, "localizations_per_s": 1e3 / ms,
           "stage_ms": {"putMatch": float(st[0]), "assembly": float(st[1]), "PnP": float(st[2])},
           "fraction_localized": ok / reps, "centre_error_m_median": float(np.median(err)) if err else None,
           "correspondences": int(len(r["corr_qfeat"])), "inliers": int(len(r["inliers"])),
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
        from oracle import oracle as orc
        orc.build()
        t0 = time.perf_counter()
        off = sc["seg_offsets"]
        m_view, m_i, m_j, m_d = [], [], [], []
        for v in range(len(off) - 1):
            oi, oj, od = orc.match_view_to_query(sc["rows"][int(off[v]):int(off[v + 1])], sc["q_desc"], 0.6)
            if len(oi) < 16:
                continue
            m_view += [v] * len(oi); m_i += oi.tolist(); m_j += oj.tolist(); m_d += od.tolist()
        t1 = time.perf_counter()
        # geometric filter of the kept views on the CPU port, timed on its own: the plain pipeline
        # below consumes the putative matches like the GPU's plain run
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
        out["cpu_baseline"] = {"ms_per_query": ((t1 - t0) + (t3 - t1g)) * 1e3, "cores": orc.num_threads(),
                               "kind": "port",
                               "stage_ms": {"putMatch": (t1 - t0) * 1e3, "assembly": (t2 - t1g) * 1e3,
                                            "PnP": (t3 - t2) * 1e3, "geoMatch_when_enabled": (t1g - t1) * 1e3},
                               "sample": "1 query: exact per-view 2-NN on all cores, sequential AC-RANSAC on one "
                                         "thread (as the reference runs it); geoMatch = F-matrix AC-RANSAC of the "
                                         "kept views one after the other on one thread, 25 rounds",
                               "localized": bool(ro["ok"]), "inliers": int(len(ro["inliers"])),
                               "geometric_pairs_valid": n_geo_valid}
        out["reference_algorithm_lsh"] = lsh_reference(sc, set(zip(m_view, m_i, m_j)))
    return out


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



def workload_config(n_gpus):
    return {"workload": "%s: %d queries x %d map descriptors (64-byte AKAZE/MLDB rows), "
                        "exact Hamming 2-NN, planted matches (30%%)"
                        % ("C5 campus-scale map" if N_MAP > 10_000_000 else "C3 building-scale map", N_QUERIES, N_MAP),
            "sharding": "map rows sharded over %d GPU(s), NCCL all-gather top-2 merge" % n_gpus if n_gpus > 1
            else "single GPU, whole table resident",
            "cache": "map table 640 MB > 126 MB L2, streamed from HBM every step (no L2 flush needed)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c3", choices=["c3", "c5"],
                    help="c3 (default, the headline): 4096 x 10M; c5: 16384 x 50M campus map (8 GPUs)")
    args = ap.parse_args()
    global N_QUERIES, N_MAP, SEED
    if args.workload == "c5":
        N_QUERIES, N_MAP, SEED = 16384, 50_000_000, 5000
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    from sfmlocalization_b200.gpu import HuloGpu, PinnedArray

    A, target, shard, row_base = make_tables(world, rank)
    g = HuloGpu(local_rank)
    id_path = None
    if world > 1:
        uid, id_path = rendezvous_id(rank, world, HuloGpu.comm_unique_id)
        g.comm_init(uid, rank, world)
    dA = g.db(A)
    dB = g.db(shard)
    nA = A.shape[0]

    def step_device():
        if world > 1:
            g.knn2_sharded(dA, dB, row_base, fetch=False)
        else:
            g.knn2(dA, dB, fetch=False)

    # ---- kernel-resident throughput: inputs already in HBM, results left in HBM
    # the clock sampler runs from the warm-up on (same load) so that short timed regions still
    # collect samples; nvidia-smi needs ~100 ms to deliver its first line
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    g.comm_barrier() if world > 1 else g.synchronize()
    launches0 = g.launch_count
    g.timer_start()
    for _ in range(args.steps):
        step_device()
    ms = g.timer_stop()
    launches = g.launch_count - launches0
    ms = g.comm_max(ms) if world > 1 else ms
    clocks = sampler.stop()
    total_dist = float(nA) * float(N_MAP)
    value = total_dist * args.steps / (ms * 1e-3) / 1e9

    # ---- end to end through the C-ABI with host buffers: every step uploads the queries from
    # pinned host memory and reads the top-2 back; the map table is engine state (resident)
    pin_A = PinnedArray(A.shape, np.uint8); pin_A.array[...] = A
    pin_i = PinnedArray((nA, 2), np.int32); pin_d = PinnedArray((nA, 2), np.int32)
    import ctypes as C
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

    step_e2e()
    g.comm_barrier() if world > 1 else g.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    g.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_s = g.comm_max(e2e_s) if world > 1 else e2e_s
    e2e_value = total_dist * args.steps / e2e_s / 1e9
    idx, dist = pin_i.array.copy(), pin_d.array.copy()

    # ---- cold variant: map shard uploaded from host inside the timed call (hulo_knn2_host)
    cold = None
    if world == 1:
        t0 = time.perf_counter()
        g.knn2_host(A, shard)
        cold_s = time.perf_counter() - t0
        cold = {"value": total_dist / cold_s / 1e9, "unit": UNIT,
                "note": "one call of hulo_knn2_host: 640 MB map + queries H2D from pageable memory inside the call"}

    # ---- sanity on the result of the timed configuration (planted rows must be found)
    ok = True
    hit = np.nonzero(target >= 0)[0]
    ok &= bool(np.array_equal(idx[hit, 0], target[hit]))
    ok &= bool((dist[:, 0] <= dist[:, 1]).all())

    if rank == 0:
        peak, peak_src = popc_peak_gdist((clocks.get("sm_max_mhz") or 1965.0))
        peaks = read_json(os.path.join(ROOT, "MEASURED_PEAKS.json"), {}) or {}
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        per_gpu = value / world
        # K1 is the only kernel of weight in the step (the merge is microseconds); its average
        # launch duration is the step time measured above with CUDA events on the library stream.
        alg_bytes = 64.0 * (nA + N_MAP / world) + 16.0 * nA
        hbm_gbs = alg_bytes / (ms * 1e-3 / args.steps) / 1e9
        traffic = (read_json(os.path.join(ROOT, "profiles", "k1_traffic.json"), {}) or {}).get("dram_bytes_per_launch")
        # the limit of the kernel's own instruction mix: per distance 26.5 LOP3 +
        # 3 VIMNMX on the ALU pipe (2 cycles per warp instruction per sub-partition) and 7.75 POPC on
        # the XU pipe (8 cycles each); the busier pipe bounds the rate
        mix_cycles = max((26.5 + 3.0) * 2.0, 7.75 * 8.0)
        mix_peak = 148 * 4 * 32 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / mix_cycles / 1e9
        roofline = {"bound": "int-popc", "achieved": per_gpu, "peak": peak, "unit": UNIT + "/GPU",
                    "frac": per_gpu / peak, "peak_source": peak_src,
                    "pipe_limit_of_kernel_mix": {"peak": mix_peak, "frac": per_gpu / mix_peak,
                                                 "cycles_per_warp_distance": mix_cycles,
                                                 "mix": "26.5 LOP3 + 3 VIMNMX (ALU, 2 clk) | 7.75 POPC (XU, 8 clk) | 7.75 IMAD (FMA)"},
                    "work_per_unit": "1 dist = 512 compared bits = 16 x 32-bit POPC (naive); the kernel folds "
                                     "words with LOP3 carry-save adders first, so frac can exceed 1",
                    "traffic": traffic,
                    "hbm": {"achieved": hbm_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_gbs / hbm_peak,
                            "algorithmic_bytes_per_launch": alg_bytes,
                            "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback"}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(nA * 64),
                    "d2h_bytes_per_step": int(nA * 16),
                    "inputs": "queries H2D from pinned memory + top-2 D2H every step; map table resident "
                              "(uploaded once at engine construction, as LocalizeEngine loads its map once)"},
            "e2e_cold": cold,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "result_check": "planted matches found, d0<=d1" if ok else "FAILED",
        }
        if world == 1 and args.workload == "c3":
            line["localize"] = localize_bench(g, with_cpu=not args.no_cpu_baseline)
        if not args.no_cpu_baseline and world == 1 and args.workload == "c3":
            cb, (ci, cd, nq, nb) = cpu_baseline_sample(A, shard)
            # the same sample through the GPU path must agree bit for bit
            gi, gd = g.knn2_host(A[:nq], shard[:nb])
            cb["gpu_equals_cpu_on_sample"] = bool(np.array_equal(gi, ci) and np.array_equal(gd, cd))
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


if __name__ == "__main__":
    main()
