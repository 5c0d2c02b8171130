"""numpy front end over the C-ABI (test / benchmark harness; the product boundary is the
C-ABI itself and the C++ hulo:: layer in csrc/host).  All arrays are host numpy arrays;
device memory lives behind the opaque handles of libhulo_gpu.so."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import HuloError, check  # noqa: F401


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _rows(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if a.ndim != 2:
        raise ValueError("descriptor rows must be a 2-D uint8 array")
    return a


class PinnedArray:
    """A numpy view over page-locked host memory (hulo_host_alloc)."""

    def __init__(self, shape, dtype):
        self.lib = _lib.load()
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape)
        n = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        check(self.lib.hulo_host_alloc(max(n, 1), C.byref(p)))
        self._p = p
        buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._p is not None:
            self.array = None
            self.lib.hulo_host_free(self._p)
            self._p = None


class DescriptorDb:
    """Device-resident 64-byte descriptor rows with a segment (view / image) table."""

    def __init__(self, gpu, rows, seg_offsets=None):
        rows = _rows(rows)
        self.gpu = gpu
        self.lib = gpu.lib
        h = C.c_void_p()
        seg = None
        n_seg = 0
        if seg_offsets is not None:
            seg = np.ascontiguousarray(seg_offsets, dtype=np.uint64)
            n_seg = len(seg) - 1
        stride = rows.shape[1] if rows.shape[0] else 64
        check(self.lib.hulo_db_upload(gpu.h, _ptr(rows), rows.shape[0], stride, _ptr(seg), n_seg, C.byref(h)))
        self.h = h

    def __len__(self):
        return int(self.lib.hulo_db_rows(self.h))

    @property
    def n_segments(self):
        return int(self.lib.hulo_db_segments(self.h))

    def update(self, rows):
        rows = _rows(rows)
        check(self.lib.hulo_db_update(self.gpu.h, self.h, _ptr(rows), rows.shape[0],
                                      rows.shape[1] if rows.shape[0] else 64))

    def download(self, first=0, n=None):
        n = len(self) - first if n is None else n
        out = np.empty((n, 64), np.uint8)
        check(self.lib.hulo_db_download(self.gpu.h, self.h, first, n, _ptr(out)))
        return out

    def free(self):
        if self.h is not None:
            self.lib.hulo_db_free(self.h)
            self.h = None


class HuloGpu:
    """One device context (hulo_gpu_create).  Raises HuloError when there is no B200."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        check(self.lib.hulo_gpu_create(device, C.byref(h)))
        self.h = h

    def close(self):
        if self.h is not None:
            self.lib.hulo_gpu_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- bookkeeping
    def db(self, rows, seg_offsets=None):
        return DescriptorDb(self, rows, seg_offsets)

    def timer_start(self):
        check(self.lib.hulo_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float(0)
        check(self.lib.hulo_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def synchronize(self):
        check(self.lib.hulo_synchronize(self.h))

    @property
    def launch_count(self):
        return int(self.lib.hulo_launch_count(self.h))

    # -- K1
    def set_knn_engine(self, engine):
        """'int' (XOR + popcount on the integer pipes), 'tc' (fp4 contraction on the tensor cores), 'tc8'
        (its int8 form) or 'auto' (the default: tc for large searches); all exact."""
        code = {"int": _lib.KNN_INT, "tc": _lib.KNN_TC, "auto": _lib.KNN_AUTO, "tc8": _lib.KNN_TC8}[engine]
        check(self.lib.hulo_gpu_set_knn_engine(self.h, code))

    @property
    def knn_engine(self):
        return {_lib.KNN_INT: "int", _lib.KNN_TC: "tc", _lib.KNN_AUTO: "auto",
                _lib.KNN_TC8: "tc8"}[int(self.lib.hulo_gpu_knn_engine(self.h))]

    def knn2(self, A, B, fetch=True):
        """A, B: DescriptorDb.  Returns (idx2, dist2) int32 nA x 2, or None when fetch=False."""
        nA = len(A)
        if not fetch:
            check(self.lib.hulo_knn2(self.h, A.h, B.h, None, None))
            return None
        idx2 = np.empty((nA, 2), np.int32)
        dist2 = np.empty((nA, 2), np.int32)
        check(self.lib.hulo_knn2(self.h, A.h, B.h, _ptr(idx2), _ptr(dist2)))
        return idx2, dist2

    def knn2_fetch(self, nA):
        idx2 = np.empty((nA, 2), np.int32)
        dist2 = np.empty((nA, 2), np.int32)
        check(self.lib.hulo_knn2_fetch(self.h, nA, _ptr(idx2), _ptr(dist2)))
        return idx2, dist2

    def knn2_host(self, A, B, idx2=None, dist2=None):
        A, B = _rows(A), _rows(B)
        nA = A.shape[0]
        idx2 = np.empty((nA, 2), np.int32) if idx2 is None else idx2
        dist2 = np.empty((nA, 2), np.int32) if dist2 is None else dist2
        check(self.lib.hulo_knn2_host(self.h, _ptr(A), nA, A.shape[1] if nA else 64, _ptr(B), B.shape[0],
                                      B.shape[1] if B.shape[0] else 64, _ptr(idx2), _ptr(dist2)))
        return idx2, dist2

    def knn2_sharded(self, A, B_shard, row_base, fetch=True):
        nA = len(A)
        if not fetch:
            check(self.lib.hulo_knn2_sharded(self.h, A.h, B_shard.h, row_base, None, None))
            return None
        idx2 = np.empty((nA, 2), np.int32)
        dist2 = np.empty((nA, 2), np.int32)
        check(self.lib.hulo_knn2_sharded(self.h, A.h, B_shard.h, row_base, _ptr(idx2), _ptr(dist2)))
        return idx2, dist2

    def knn2_sharded_submit(self, A, B_shard, row_base):
        """Pipelined form: issue the search and return at once (see knn2_sharded_collect)."""
        check(self.lib.hulo_knn2_sharded_submit(self.h, A.h, B_shard.h, row_base))
        self._pipe_rows = getattr(self, "_pipe_rows", [])
        self._pipe_rows.append(len(A))

    def knn2_sharded_collect(self, idx2=None, dist2=None):
        """Results of the oldest submitted search not collected yet -> (idx2, dist2)."""
        nA = self._pipe_rows.pop(0)
        idx2 = np.empty((nA, 2), np.int32) if idx2 is None else idx2
        dist2 = np.empty((nA, 2), np.int32) if dist2 is None else dist2
        n = C.c_size_t(0)
        check(self.lib.hulo_knn2_sharded_collect(self.h, _ptr(idx2), _ptr(dist2), C.byref(n)))
        assert n.value == nA
        return idx2, dist2

    def merge_top2(self, cand):
        """cand: world x nA x 4 int32 records {d0, i0, d1, i1} with global indices."""
        cand = np.ascontiguousarray(cand, np.int32)
        world, nA = cand.shape[0], cand.shape[1]
        idx2 = np.empty((nA, 2), np.int32)
        dist2 = np.empty((nA, 2), np.int32)
        check(self.lib.hulo_merge_top2(self.h, _ptr(cand), nA, world, _ptr(idx2), _ptr(dist2)))
        return idx2, dist2

    def match_to_query(self, map_db, query, ratio, views=None, cap=None):
        """hulo_match_to_query -> dict(view, i, j, d0, view_counts)."""
        query = _rows(query)
        nq = query.shape[0]
        if views is not None:
            views = np.ascontiguousarray(views, dtype=np.uint32)
            n_views = len(views)
        else:
            n_views = map_db.n_segments
        if cap is None:
            cap = max(len(map_db), 1)
        ov = np.empty(cap, np.uint32); oi = np.empty(cap, np.uint32)
        oj = np.empty(cap, np.uint32); od = np.empty(cap, np.int32)
        vc = np.zeros(max(n_views, 1), np.uint32)
        n = C.c_size_t(0)
        check(self.lib.hulo_match_to_query(self.h, map_db.h, _ptr(views), n_views, _ptr(query), nq,
                                           query.shape[1] if nq else 64, ratio, _ptr(ov), _ptr(oi), _ptr(oj),
                                           _ptr(od), cap, C.byref(n), _ptr(vc)))
        k = n.value
        return dict(view=ov[:k].copy(), i=oi[:k].copy(), j=oj[:k].copy(), d0=od[:k].copy(),
                    view_counts=vc[:n_views].copy())

    def match_to_queries(self, map_db, queries, q_offsets, ratio, views=None, cap=None):
        """hulo_match_to_queries -> dict(query, view, i, j, d0, counts[n_queries, n_views])."""
        queries = _rows(queries)
        q_offsets = np.ascontiguousarray(q_offsets, np.uint64)
        nQ = len(q_offsets) - 1
        if views is not None:
            views = np.ascontiguousarray(views, dtype=np.uint32)
            n_views = len(views)
        else:
            n_views = map_db.n_segments
        cap = (1 << 20) if cap is None else cap
        while True:
            oq = np.empty(cap, np.uint32); ov = np.empty(cap, np.uint32); oi = np.empty(cap, np.uint32)
            oj = np.empty(cap, np.uint32); od = np.empty(cap, np.int32)
            counts = np.zeros((max(nQ, 1), max(n_views, 1)), np.uint32)
            n = C.c_size_t(0)
            st = self.lib.hulo_match_to_queries(self.h, map_db.h, _ptr(views), n_views, _ptr(queries),
                                                queries.shape[1] if queries.shape[0] else 64, _ptr(q_offsets), nQ,
                                                ratio, _ptr(oq), _ptr(ov), _ptr(oi), _ptr(oj), _ptr(od), cap,
                                                C.byref(n), _ptr(counts))
            if st == _lib.ERR_CAPACITY:
                cap = int(n.value)
                continue
            check(st)
            k = n.value
            return dict(query=oq[:k].copy(), view=ov[:k].copy(), i=oi[:k].copy(), j=oj[:k].copy(), d0=od[:k].copy(),
                        counts=counts[:nQ, :n_views].copy())

    def match_pairs(self, db, pairs, ratio, flags=_lib.PAIR_REFERENCE, cap=None):
        """hulo_match_pairs -> (pair_offsets uint64[P+1], i, j)."""
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        P = pairs.shape[0]
        if cap is None:
            cap = 1 << 20
        while True:
            off = np.zeros(P + 1, np.uint64)
            oi = np.empty(max(cap, 1), np.uint32)
            oj = np.empty(max(cap, 1), np.uint32)
            n = C.c_size_t(0)
            st = self.lib.hulo_match_pairs(self.h, db.h, _ptr(pairs), P, ratio, flags, _ptr(off), _ptr(oi),
                                           _ptr(oj), cap, C.byref(n))
            if st == _lib.ERR_CAPACITY:
                cap = int(n.value)
                continue
            check(st)
            return off, oi[:n.value].copy(), oj[:n.value].copy()

    # -- K2
    def score_resection(self, models, x2d, X3d, K, thr_px=-1.0):
        models = np.ascontiguousarray(models, np.float64).reshape(-1, 12)
        x2d = np.ascontiguousarray(x2d, np.float64); X3d = np.ascontiguousarray(X3d, np.float64)
        K = np.ascontiguousarray(K, np.float64)
        H, N = models.shape[0], x2d.shape[0]
        nfa = np.empty(H, np.float32); kb = np.empty(H, np.int32)
        ek = np.empty(H, np.float32); ni = np.empty(H, np.int32)
        check(self.lib.hulo_score_resection(self.h, _ptr(models), H, _ptr(x2d), _ptr(X3d), N, _ptr(K), thr_px,
                                            _ptr(nfa), _ptr(kb), _ptr(ek), _ptr(ni)))
        return nfa, kb, ek, ni

    def resection_residuals(self, models, x2d, X3d, K):
        models = np.ascontiguousarray(models, np.float64).reshape(-1, 12)
        x2d = np.ascontiguousarray(x2d, np.float64); X3d = np.ascontiguousarray(X3d, np.float64)
        K = np.ascontiguousarray(K, np.float64)
        H, N = models.shape[0], x2d.shape[0]
        res = np.empty((H, N), np.float32)
        check(self.lib.hulo_resection_residuals(self.h, _ptr(models), H, _ptr(x2d), _ptr(X3d), N, _ptr(K),
                                                _ptr(res)))
        return res

    def p3p(self, triplets, x2d, X3d, K):
        triplets = np.ascontiguousarray(triplets, np.uint32).reshape(-1, 3)
        x2d = np.ascontiguousarray(x2d, np.float64); X3d = np.ascontiguousarray(X3d, np.float64)
        K = np.ascontiguousarray(K, np.float64)
        T = triplets.shape[0]
        models = np.zeros((T, 4, 3, 4)); nm = np.zeros(T, np.int32)
        check(self.lib.hulo_p3p(self.h, _ptr(triplets), T, _ptr(x2d), _ptr(X3d), x2d.shape[0], _ptr(K),
                                _ptr(models), _ptr(nm)))
        return models, nm

    def resect_acransac(self, x2d, X3d, K, max_iter=4096, seed=1, sequential=False):
        x2d = np.ascontiguousarray(x2d, np.float64); X3d = np.ascontiguousarray(X3d, np.float64)
        K = np.ascontiguousarray(K, np.float64)
        N = x2d.shape[0]
        P = np.zeros((3, 4)); inl = np.empty(max(N, 1), np.int32)
        n_inl = C.c_size_t(0); emax = C.c_double(0); found = C.c_int(0)
        fn = self.lib.hulo_resect_acransac_sequential if sequential else self.lib.hulo_resect_acransac
        check(fn(self.h, _ptr(x2d), _ptr(X3d), N, _ptr(K), max_iter, seed, _ptr(P),
                                            _ptr(inl), C.byref(n_inl), C.byref(emax), C.byref(found)))
        return dict(found=bool(found.value), P=P, inliers=inl[:n_inl.value].copy(), error_max=emax.value)

    def pose_from_projection(self, P):
        """hulo_pose_from_projection -> (K, R, center)."""
        P = np.ascontiguousarray(P, np.float64)
        K = np.zeros((3, 3)); R = np.zeros((3, 3)); c = np.zeros(3)
        check(self.lib.hulo_pose_from_projection(_ptr(P), _ptr(K), _ptr(R), _ptr(c)))
        return K, R, c

    def resect_acransac_batch(self, x2d, X3d, offsets, K, max_iter=4096, seed=1, seeds=None):
        """hulo_resect_acransac_batch -> list of dict(found, P, inliers, error_max), one per problem.
        K: (n, 3, 3) or one (3, 3) shared by all problems."""
        x2d = np.ascontiguousarray(x2d, np.float64).reshape(-1, 2)
        X3d = np.ascontiguousarray(X3d, np.float64).reshape(-1, 3)
        off = np.ascontiguousarray(offsets, np.uint64)
        n = len(off) - 1
        K = np.asarray(K, np.float64)
        K = np.ascontiguousarray(np.broadcast_to(K.reshape(-1, 3, 3), (max(n, 1), 3, 3)))
        if seeds is not None:
            seeds = np.ascontiguousarray(seeds, np.uint64)
        total = int(off[-1]) if n > 0 else 0
        P = np.zeros((max(n, 1), 3, 4)); inl = np.empty(max(total, 1), np.int32)
        ninl = np.zeros(max(n, 1), np.uint64); emax = np.zeros(max(n, 1)); found = np.zeros(max(n, 1), np.int32)
        check(self.lib.hulo_resect_acransac_batch(self.h, n, _ptr(off), _ptr(x2d), _ptr(X3d), _ptr(K), max_iter, seed,
                                                  _ptr(seeds), _ptr(P), _ptr(inl), _ptr(ninl), _ptr(emax),
                                                  _ptr(found)))
        return [dict(found=bool(found[p]), P=P[p].copy(),
                     inliers=inl[int(off[p]):int(off[p]) + int(ninl[p])].copy(), error_max=float(emax[p]))
                for p in range(n)]

    # -- K3
    def geometric_filter(self, xI, xJ, pair_offsets, image_sizes, precision_px=4.0, max_iter=25, seed=1,
                         pair_seeds=None):
        """hulo_geometric_filter -> dict(valid, n_inliers, inliers (list per pair), F, error_max, nfa)."""
        xI = np.ascontiguousarray(xI, np.float64).reshape(-1, 2)
        xJ = np.ascontiguousarray(xJ, np.float64).reshape(-1, 2)
        off = np.ascontiguousarray(pair_offsets, np.uint64)
        P = len(off) - 1
        sizes = np.ascontiguousarray(image_sizes, np.int32).reshape(-1, 4)
        total = int(off[-1]) if P >= 0 and len(off) else 0
        if pair_seeds is not None:
            pair_seeds = np.ascontiguousarray(pair_seeds, np.uint64)
        valid = np.zeros(max(P, 1), np.int32); ninl = np.zeros(max(P, 1), np.uint32)
        inl = np.zeros(max(total, 1), np.int32)
        F = np.zeros((max(P, 1), 3, 3)); emax = np.zeros(max(P, 1)); nfa = np.zeros(max(P, 1))
        check(self.lib.hulo_geometric_filter(self.h, _ptr(xI), _ptr(xJ), _ptr(off), P, _ptr(sizes), precision_px,
                                             max_iter, seed, _ptr(pair_seeds), _ptr(valid), _ptr(ninl), _ptr(inl),
                                             _ptr(F), _ptr(emax), _ptr(nfa)))
        per_pair = [inl[int(off[p]):int(off[p]) + int(ninl[p])].copy() for p in range(P)]
        return dict(valid=valid[:P].astype(bool), n_inliers=ninl[:P].copy(), inliers=per_pair, F=F[:P].copy(),
                    error_max=emax[:P].copy(), nfa=nfa[:P].copy())

    def guided_match(self, db, xy, pairs, F, error_th, dist_ratio=0.36, dedup=True, cap=None):
        """hulo_guided_match -> (pair_offsets uint64[P+1], i, j)."""
        xy = np.ascontiguousarray(xy, np.float64).reshape(-1, 2)
        pairs = np.ascontiguousarray(pairs, np.uint32).reshape(-1, 2)
        F = np.ascontiguousarray(F, np.float64).reshape(-1, 9)
        thr = np.ascontiguousarray(error_th, np.float64)
        P = pairs.shape[0]
        cap = (1 << 20) if cap is None else cap
        while True:
            off = np.zeros(P + 1, np.uint64)
            oi = np.empty(max(cap, 1), np.uint32)
            oj = np.empty(max(cap, 1), np.uint32)
            n = C.c_size_t(0)
            st = self.lib.hulo_guided_match(self.h, db.h, _ptr(xy), _ptr(pairs), P, _ptr(F), _ptr(thr), dist_ratio,
                                            int(bool(dedup)), _ptr(off), _ptr(oi), _ptr(oj), cap, C.byref(n))
            if st == _lib.ERR_CAPACITY:
                cap = int(n.value)
                continue
            check(st)
            return off, oi[:n.value].copy(), oj[:n.value].copy()

    # -- K4
    def ransac_transform3d(self, A, B, thres, rounds, svd_ratio=float("inf"), similarity=False, samples=None, seed=1):
        """hulo_ransac_transform3d -> (M 3x4, inliers) or (empty, empty) like the reference."""
        A = np.ascontiguousarray(A, np.float64); B = np.ascontiguousarray(B, np.float64)
        n = A.shape[1]
        if samples is not None:
            samples = np.ascontiguousarray(samples, np.uint32).reshape(-1, 4)
            rounds = samples.shape[0]
        M = np.zeros((3, 4)); inl = np.zeros(max(n, 1), np.int32)
        ni = C.c_size_t(0); br = C.c_uint32(0)
        check(self.lib.hulo_ransac_transform3d(self.h, _ptr(A), _ptr(B), n, thres, _ptr(samples), rounds, seed,
                                               svd_ratio, int(bool(similarity)), _ptr(M), _ptr(inl), C.byref(ni),
                                               C.byref(br)))
        if ni.value == 0:
            return np.array([]), np.asarray([], np.int64)
        return M, inl[:ni.value].astype(np.int64)

    # -- multi GPU
    @staticmethod
    def comm_unique_id():
        buf = (C.c_uint8 * 128)()
        check(_lib.load().hulo_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, unique_id, rank, world):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        check(self.lib.hulo_comm_init(self.h, buf, rank, world))

    @property
    def exchange_kind(self):
        """Data plane of the row-sharded search on this context: 'peer-store', 'nccl-allgather' or 'none'."""
        return self.lib.hulo_comm_exchange_kind(self.h).decode()

    def comm_barrier(self):
        check(self.lib.hulo_comm_barrier(self.h))

    def comm_max(self, value):
        v = C.c_double(value)
        check(self.lib.hulo_comm_max_f64(self.h, C.byref(v)))
        return v.value


class BowIndex:
    """hulo_bow_*: the BoF vectors of the map views, resident on the device (hulo::selectViewByBoF)."""

    def __init__(self, gpu, bof):
        bof = np.ascontiguousarray(bof, np.float32)
        self.lib = gpu.lib
        h = C.c_void_p()
        check(self.lib.hulo_bow_create(gpu.h, _ptr(bof), bof.shape[0], bof.shape[1], C.byref(h)))
        self.h = h

    def knn(self, query, knn, subset=None):
        query = np.ascontiguousarray(query, np.float32).ravel()
        ns = 0
        if subset is not None:
            subset = np.ascontiguousarray(subset, np.uint32)
            ns = len(subset)
        idx = np.zeros(max(knn, 1), np.int32); dist = np.zeros(max(knn, 1), np.float32)
        check(self.lib.hulo_bow_knn(self.h, _ptr(query), _ptr(subset), ns, knn, _ptr(idx), _ptr(dist)))
        return idx[:knn].copy(), dist[:knn].copy()

    def close(self):
        if self.h is not None:
            self.lib.hulo_bow_destroy(self.h)
            self.h = None


class LocalizeEngine:
    """hulo_engine_*: the hot path of LocalizeEngine::localize (LocalizeEngine.cc:423-602) with
    the map resident on the device."""

    def __init__(self, gpu, rows, seg_offsets, obs_view, obs_feat, obs_landmark, landmark_X, K,
                 ratio=0.6, min_putative=16, min_points=8, min_inliers=10, max_iter=4096):
        self.gpu = gpu
        self.lib = gpu.lib
        rows = _rows(rows)
        seg = np.ascontiguousarray(seg_offsets, np.uint64)
        ov = np.ascontiguousarray(obs_view, np.uint32); of = np.ascontiguousarray(obs_feat, np.uint32)
        ol = np.ascontiguousarray(obs_landmark, np.uint32)
        X = np.ascontiguousarray(landmark_X, np.float64); K = np.ascontiguousarray(K, np.float64)
        h = C.c_void_p()
        check(self.lib.hulo_engine_create(gpu.h, _ptr(rows), rows.shape[0], rows.shape[1] if rows.shape[0] else 64,
                                          _ptr(seg), len(seg) - 1, _ptr(ov), _ptr(of), _ptr(ol), len(ov), _ptr(X),
                                          X.shape[0], _ptr(K), C.byref(h)))
        self.h = h
        check(self.lib.hulo_engine_configure(self.h, ratio, min_putative, min_points, min_inliers, max_iter))

    def set_keypoints(self, map_xy, view_wh, query_wh):
        map_xy = np.ascontiguousarray(map_xy, np.float64)
        view_wh = np.ascontiguousarray(view_wh, np.int32)
        check(self.lib.hulo_engine_set_keypoints(self.h, _ptr(map_xy), _ptr(view_wh), int(query_wh[0]),
                                                 int(query_wh[1])))

    def set_resection_schedule(self, schedule):
        """'batched' (default) or 'sequential' (the reference's AC-RANSAC loop kept to the letter)."""
        check(self.lib.hulo_engine_set_resection_schedule(self.h, {"batched": 0, "sequential": 1}[schedule]))

    def configure_geometric(self, enabled, ransac_round=25, precision_px=4.0):
        check(self.lib.hulo_engine_configure_geometric(self.h, int(bool(enabled)), ransac_round, precision_px))

    def set_guided_matching(self, enabled):
        check(self.lib.hulo_engine_set_guided_matching(self.h, int(bool(enabled))))

    def localize_sharded(self, qdesc, qxy, views=None, seed=1):
        """hulo_engine_localize_sharded: collective over the ranks of the context's communicator."""
        return self.localize(qdesc, qxy, views, seed, _fn=self.lib.hulo_engine_localize_sharded)

    def localize(self, qdesc, qxy, views=None, seed=1, _fn=None):
        qdesc = _rows(qdesc)
        qxy = np.ascontiguousarray(qxy, np.float64)
        nq = qdesc.shape[0]
        n_views = 0
        if views is not None:
            views = np.ascontiguousarray(views, np.uint32)
            n_views = len(views)
        pose = np.zeros(12); loc = C.c_int(0)
        cq = np.empty(max(nq, 1), np.uint32); cl = np.empty(max(nq, 1), np.uint32)
        inl = np.empty(max(nq, 1), np.int32)
        nc = C.c_size_t(0); ni = C.c_size_t(0)
        times = np.zeros(4)
        fn = _fn if _fn is not None else self.lib.hulo_engine_localize
        check(fn(self.h, _ptr(qdesc), nq, qdesc.shape[1] if nq else 64, _ptr(qxy), _ptr(views), n_views, seed, _ptr(pose),
                 C.byref(loc), _ptr(cq), _ptr(cl), C.byref(nc), _ptr(inl), C.byref(ni), _ptr(times)))
        return dict(localized=bool(loc.value), center=pose[:3].copy(), R=pose[3:].reshape(3, 3).copy(),
                    corr_qfeat=cq[:nc.value].copy(), corr_landmark=cl[:nc.value].copy(),
                    inliers=inl[:ni.value].copy(), times_ms=times)

    def localize_batch(self, qdescs, qxys, views=None, seed=1):
        """qdescs / qxys: lists of per-image arrays.  Returns dict of per-image arrays."""
        nQ = len(qdescs)
        off = np.zeros(nQ + 1, np.uint64)
        off[1:] = np.cumsum([d.shape[0] for d in qdescs])
        desc = _rows(np.concatenate(qdescs, axis=0)) if nQ else np.zeros((0, 64), np.uint8)
        xy = np.ascontiguousarray(np.concatenate(qxys, axis=0), np.float64) if nQ else np.zeros((0, 2))
        n_views = 0
        if views is not None:
            views = np.ascontiguousarray(views, np.uint32)
            n_views = len(views)
        pose = np.zeros((max(nQ, 1), 12)); loc = np.zeros(max(nQ, 1), np.int32)
        nc = np.zeros(max(nQ, 1), np.uint32); ni = np.zeros(max(nQ, 1), np.uint32)
        times = np.zeros(4)
        check(self.lib.hulo_engine_localize_batch(self.h, nQ, _ptr(desc), desc.shape[1] if desc.shape[0] else 64,
                                                  _ptr(off), _ptr(xy), _ptr(views), n_views, seed, _ptr(pose),
                                                  _ptr(loc), _ptr(nc), _ptr(ni), _ptr(times)))
        return dict(localized=loc[:nQ].astype(bool), center=pose[:nQ, :3].copy(),
                    R=pose[:nQ, 3:].reshape(-1, 3, 3).copy(), n_corr=nc[:nQ].copy(), n_inliers=ni[:nQ].copy(),
                    times_ms=times)

    def close(self):
        if self.h is not None:
            self.lib.hulo_engine_destroy(self.h)
            self.h = None


# ---- the reference's Python entry points for the model-merge RANSAC, same names and signatures
# (PyVisionLocalizeCommon/src/hulo_sfm/mergeSfM.py:344, 394; hulo_transform/ransacTransform.py:13),
# on the GPU.  They use one lazily created context on device 0.
_default_gpu = None


def _gpu0():
    global _default_gpu
    if _default_gpu is None:
        _default_gpu = HuloGpu(0)
    return _default_gpu


def ransacAffineTransform(A, B, thres, ransacRound, svdRatio=float("inf"), seed=1):
    import sys
    return _gpu0().ransac_transform3d(A, B, thres, ransacRound, min(svdRatio, sys.float_info.max), False, seed=seed)


def ransacSimilarityTransform(A, B, thres, ransacRound, svdRatio=float("inf"), seed=1):
    import sys
    return _gpu0().ransac_transform3d(A, B, thres, ransacRound, min(svdRatio, sys.float_info.max), True, seed=seed)


def ransacTransform(A, B, thres, ransacRound, svdRatio=float("inf"), seed=1):
    """mergeSfM.ransacTransform: the similarity flavour when hulo_transform is importable (it is part
    of the reference tree), which is what the drivers get."""
    return ransacSimilarityTransform(A, B, thres, ransacRound, svdRatio, seed)


def partition_views(rows_per_view, world):
    """hulo_partition_views -> bounds (world + 1): rank r matches views[bounds[r]:bounds[r + 1]]."""
    rows = np.ascontiguousarray(rows_per_view, np.uint64)
    bounds = np.zeros(world + 1, np.uint64)
    check(_lib.load().hulo_partition_views(_ptr(rows), len(rows), world, _ptr(bounds)))
    return bounds.astype(np.int64)
