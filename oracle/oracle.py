"""ctypes/numpy front end of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
Each function documents the C function it wraps; those cite the reference lines.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
INT_MAX = 2**31 - 1


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc, OpenMP)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle_match.c", "oracle_resect.c", "Makefile")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs)):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_logcombi.restype = C.c_float
        _lib.orc_logcombi.argtypes = [C.c_size_t, C.c_size_t]
        _lib.orc_best_nfa.restype = C.c_double
        _lib.orc_ratio_pass.argtypes = [C.c_int32, C.c_int32, C.c_float]
        for name in ("orc_match_view_to_query", "orc_pair_filter", "orc_match_pair",
                     "orc_track_propagate", "orc_match_set"):
            getattr(_lib, name).restype = C.c_size_t
    return _lib


def _p(a, t=None):
    return a.ctypes.data_as(C.c_void_p)


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2
    return a


def num_threads():
    return int(lib().orc_num_threads())


def use_all_cores():
    """Run the OpenMP loops on every core this process may use, ignoring OMP_NUM_THREADS
    (torchrun exports OMP_NUM_THREADS=1 to its ranks).  Returns the thread count."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().orc_set_num_threads(int(n))
    return num_threads()


def pad_rows(rows):
    """orc_pad_rows: N x w (w <= 64) -> N x 64, zero padded (FileUtils.cpp:77-92)."""
    rows = _u8(rows)
    out = np.empty((rows.shape[0], 64), np.uint8)
    lib().orc_pad_rows(_p(rows), C.c_size_t(rows.shape[0]), C.c_size_t(rows.shape[1]), _p(out))
    return out


def knn2(A, B):
    """orc_knn2_hamming: exact 2-NN, (dist, idx) order.  Returns idx2, dist2 (nA x 2 int32)."""
    A, B = _u8(A), _u8(B)
    nA = A.shape[0]
    idx2 = np.empty((nA, 2), np.int32)
    dist2 = np.empty((nA, 2), np.int32)
    lib().orc_knn2_hamming(_p(A), C.c_size_t(nA), C.c_size_t(A.shape[1]), _p(B),
                           C.c_size_t(B.shape[0]), C.c_size_t(B.shape[1] if B.ndim == 2 else 64),
                           _p(idx2), _p(dist2))
    return idx2, dist2


def ratio_pass(d0, d1, ratio):
    return bool(lib().orc_ratio_pass(int(d0), int(d1), float(ratio)))


def match_view_to_query(A, Bq, ratio):
    """orc_match_view_to_query (MatchUtils.cpp:339-355) -> (i, j, d0) arrays."""
    A, Bq = _u8(A), _u8(Bq)
    nA = A.shape[0]
    oi = np.empty(max(nA, 1), np.int32)
    oj = np.empty(max(nA, 1), np.int32)
    od = np.empty(max(nA, 1), np.int32)
    n = lib().orc_match_view_to_query(_p(A), C.c_size_t(nA), C.c_size_t(A.shape[1]), _p(Bq),
                                      C.c_size_t(Bq.shape[0]), C.c_size_t(Bq.shape[1]),
                                      C.c_float(ratio), _p(oi), _p(oj), _p(od))
    return oi[:n].copy(), oj[:n].copy(), od[:n].copy()


def pair_filter(idx2, dist2, ratio):
    """orc_pair_filter (MatchUtils.cpp:111-150) -> (i, j) arrays."""
    idx2 = np.ascontiguousarray(idx2, np.int32)
    dist2 = np.ascontiguousarray(dist2, np.int32)
    nA = idx2.shape[0]
    oi = np.empty(max(nA, 1), np.int32)
    oj = np.empty(max(nA, 1), np.int32)
    n = lib().orc_pair_filter(_p(idx2), _p(dist2), C.c_size_t(nA), C.c_float(ratio), _p(oi), _p(oj))
    return oi[:n].copy(), oj[:n].copy()


def match_pair(A, B, ratio):
    """orc_match_pair (MatchUtils.cpp:94-150) -> (i, j) arrays."""
    A, B = _u8(A), _u8(B)
    nA = A.shape[0]
    oi = np.empty(max(nA, 1), np.int32)
    oj = np.empty(max(nA, 1), np.int32)
    n = lib().orc_match_pair(_p(A), C.c_size_t(nA), C.c_size_t(A.shape[1]), _p(B),
                             C.c_size_t(B.shape[0]), C.c_size_t(B.shape[1]), C.c_float(ratio),
                             _p(oi), _p(oj))
    return oi[:n].copy(), oj[:n].copy()


def track_propagate(n_frames, max_frame_dist, feat_number, m_off, m_i, m_j):
    """orc_track_propagate (MatchUtils.cpp:239-276) -> (f, to, i, j) arrays."""
    feat_number = np.ascontiguousarray(feat_number, np.int32)
    m_off = np.ascontiguousarray(m_off, np.int64)
    m_i = np.ascontiguousarray(m_i, np.int32)
    m_j = np.ascontiguousarray(m_j, np.int32)
    need = C.c_size_t(0)
    dummy = np.empty(1, np.int32)
    lib().orc_track_propagate(C.c_size_t(n_frames), C.c_size_t(max_frame_dist), _p(feat_number),
                              _p(m_off), _p(m_i), _p(m_j), _p(dummy), _p(dummy), _p(dummy),
                              _p(dummy), C.c_size_t(0), C.byref(need))
    cap = max(int(need.value), 1)
    of, ot, oi, oj = (np.empty(cap, np.int32) for _ in range(4))
    n = lib().orc_track_propagate(C.c_size_t(n_frames), C.c_size_t(max_frame_dist), _p(feat_number),
                                  _p(m_off), _p(m_i), _p(m_j), _p(of), _p(ot), _p(oi), _p(oj),
                                  C.c_size_t(cap), C.byref(need))
    return of[:n].copy(), ot[:n].copy(), oi[:n].copy(), oj[:n].copy()


def match_set(m_view, m_i, m_j, fd_view, fd_j, fd_d, lm_view, lm_feat, lm_id, n_query_feat):
    """orc_match_set (SfMDataUtils.cpp:59-125) -> (query feat, landmark id) arrays."""
    a32 = lambda x: np.ascontiguousarray(x, np.int32)
    m_view, m_i, m_j = a32(m_view), a32(m_i), a32(m_j)
    fd_view, fd_j, fd_d = a32(fd_view), a32(fd_j), a32(fd_d)
    lm_view, lm_feat = a32(lm_view), a32(lm_feat)
    lm_id = np.ascontiguousarray(lm_id, np.int64)
    oj = np.empty(max(n_query_feat, 1), np.int32)
    ol = np.empty(max(n_query_feat, 1), np.int64)
    n = lib().orc_match_set(_p(m_view), _p(m_i), _p(m_j), C.c_size_t(len(m_view)),
                            _p(fd_view), _p(fd_j), _p(fd_d), C.c_size_t(len(fd_view)),
                            _p(lm_view), _p(lm_feat), _p(lm_id), C.c_size_t(len(lm_view)),
                            C.c_size_t(n_query_feat), _p(oj), _p(ol))
    return oj[:n].copy(), ol[:n].copy()


# ---------------------------------------------------------------- resection
def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def p3p(bearings, X):
    """orc_p3p: bearings 3x3 (rows = unit vectors), X 3x3 (rows = points) -> list of 3x4."""
    f, X = _f64(bearings), _f64(X)
    out = np.zeros((4, 3, 4))
    n = lib().orc_p3p(_p(f), _p(X), _p(out))
    return out[:n].copy()


def normalize_points(x2d, K):
    """x2d N x 2 pixels -> N x 2 normalised (K^-1)."""
    x2d, K = _f64(x2d), _f64(K)
    out = np.empty_like(x2d)
    lib().orc_normalize_points(_p(x2d), C.c_size_t(x2d.shape[0]), _p(K), _p(out))
    return out


def residuals(M, x2dn, X3d):
    M, x2dn, X3d = _f64(M), _f64(x2dn), _f64(X3d)
    N = x2dn.shape[0]
    err = np.empty(N)
    lib().orc_residuals(_p(M), _p(x2dn), _p(X3d), C.c_size_t(N), _p(err))
    return err


def logcombi(k, n):
    return float(lib().orc_logcombi(k, n))


def best_nfa(sorted_err, N=None):
    e = _f64(sorted_err)
    N = len(e) if N is None else N
    lcn = np.empty(N + 1, np.float32)
    lck = np.empty(N + 1, np.float32)
    lib().orc_make_logcombi(C.c_size_t(N), _p(lcn), _p(lck))
    kb = C.c_size_t(0)
    nfa = lib().orc_best_nfa(_p(e), C.c_size_t(N), C.c_double(np.log10(np.pi)),
                             C.c_double(np.log10(4.0 * (N - 3))), C.c_double(np.inf),
                             _p(lcn), _p(lck), C.c_double(1.0), C.byref(kb))
    return float(nfa), int(kb.value)


def score_hypotheses(models, x2dn, X3d, thr2=-1.0):
    """orc_score_hypotheses: models H x 3 x 4 -> nfa[H], k_best[H], err_k[H], n_inl[H]."""
    models, x2dn, X3d = _f64(models), _f64(x2dn), _f64(X3d)
    H = models.shape[0]
    N = x2dn.shape[0]
    nfa = np.empty(H)
    kb = np.empty(H, np.int32)
    ek = np.empty(H)
    ni = np.empty(H, np.int32)
    lib().orc_score_hypotheses(_p(models), C.c_size_t(H), _p(x2dn), _p(X3d), C.c_size_t(N),
                               C.c_double(thr2), _p(nfa), _p(kb), _p(ek), _p(ni))
    return nfa, kb, ek, ni


def acransac(x2d, X3d, K, max_iter=4096, seed=1):
    """orc_acransac -> dict(ok, P 3x4, inliers, error_max [px], nfa)."""
    x2d, X3d, K = _f64(x2d), _f64(X3d), _f64(K)
    N = x2d.shape[0]
    P = np.zeros((3, 4))
    inl = np.empty(max(N, 1), np.int32)
    n_inl = C.c_size_t(0)
    emax = C.c_double(0)
    nfa = C.c_double(0)
    ok = lib().orc_acransac(_p(x2d), _p(X3d), C.c_size_t(N), _p(K), C.c_size_t(max_iter),
                            C.c_uint64(seed), _p(P), _p(inl), C.byref(n_inl), C.byref(emax),
                            C.byref(nfa))
    return dict(ok=bool(ok), P=P, inliers=inl[:n_inl.value].copy(), error_max=emax.value,
                nfa=nfa.value)


def krt_from_p(P):
    P = _f64(P)
    K = np.empty((3, 3)); R = np.empty((3, 3)); t = np.empty(3); c = np.empty(3)
    lib().orc_krt_from_p(_p(P), _p(K), _p(R), _p(t), _p(c))
    return K, R, t, c


# ---------------------------------------------------------------- F-matrix geometric filter
def precondition(w, h):
    T = np.empty((3, 3))
    lib().orc_precondition(int(w), int(h), _p(T))
    return T


def seven_point(x1, x2):
    """orc_seven_point: 7 x 2 normalised points each -> up to three 3x3 F (x2^T F x1 = 0)."""
    x1, x2 = _f64(x1), _f64(x2)
    F = np.zeros((3, 3, 3))
    n = lib().orc_seven_point(_p(x1), _p(x2), _p(F))
    return F[:n].copy()


def epipolar_errors(F, x1, x2):
    F, x1, x2 = _f64(F), _f64(x1), _f64(x2)
    err = np.empty(x1.shape[0])
    lib().orc_epipolar_errors(_p(F), _p(x1), _p(x2), C.c_size_t(x1.shape[0]), _p(err))
    return err


def fmatrix_score(F, x1n, x2n, logalpha0, max_thr=np.inf):
    F, x1n, x2n = _f64(F), _f64(x1n), _f64(x2n)
    kb = C.c_size_t(0)
    ek = C.c_double(0)
    lib().orc_fmatrix_score.restype = C.c_double
    nfa = lib().orc_fmatrix_score(_p(F), _p(x1n), _p(x2n), C.c_size_t(x1n.shape[0]), C.c_double(logalpha0),
                                  C.c_double(max_thr), C.byref(kb), C.byref(ek))
    return float(nfa), int(kb.value), float(ek.value)


def fmatrix_acransac(xI, xJ, sizeI, sizeJ, precision_px=4.0, max_iter=1024, seed=1):
    """orc_fmatrix_acransac -> dict(ok, F, inliers, error_max [px], nfa)."""
    xI, xJ = _f64(xI), _f64(xJ)
    N = xI.shape[0]
    F = np.zeros((3, 3))
    inl = np.empty(max(N, 1), np.int32)
    n_inl = C.c_size_t(0)
    emax = C.c_double(0)
    nfa = C.c_double(0)
    ok = lib().orc_fmatrix_acransac(_p(xI), _p(xJ), C.c_size_t(N), int(sizeI[0]), int(sizeI[1]), int(sizeJ[0]),
                                    int(sizeJ[1]), C.c_double(precision_px), C.c_size_t(max_iter), C.c_uint64(seed),
                                    _p(F), _p(inl), C.byref(n_inl), C.byref(emax), C.byref(nfa))
    return dict(ok=bool(ok), F=F, inliers=inl[:n_inl.value].copy(), error_max=emax.value, nfa=nfa.value)


# ---------------------------------------------------------------- guided matching (geometric filter, -gm)
def guided_match(F, xI, descI, xJ, descJ, error_th, dist_ratio=0.36, dedup=True):
    """orc_guided_match (+ orc_guided_dedup) -> (i, j) arrays, ascending i."""
    F, xI, xJ = _f64(F), _f64(xI), _f64(xJ)
    descI, descJ = _u8(descI), _u8(descJ)
    nI, nJ = descI.shape[0], descJ.shape[0]
    oi = np.empty(max(nI, 1), np.int32)
    oj = np.empty(max(nI, 1), np.int32)
    lib().orc_guided_match.restype = C.c_size_t
    lib().orc_guided_dedup.restype = C.c_size_t
    n = lib().orc_guided_match(_p(F), _p(xI), _p(descI), C.c_size_t(nI), C.c_size_t(descI.shape[1] if nI else 64),
                               _p(xJ), _p(descJ), C.c_size_t(nJ), C.c_size_t(descJ.shape[1] if nJ else 64),
                               C.c_double(error_th), C.c_double(dist_ratio), _p(oi), _p(oj))
    if dedup and n:
        n = lib().orc_guided_dedup(_p(xI), _p(xJ), _p(oi), _p(oj), C.c_size_t(n))
    return oi[:n].copy(), oj[:n].copy()


# ---------------------------------------------------------------- 3D-3D model-merge RANSAC
def superimposition_matrix(v0, v1):
    """Restatement of transformations.superimposition_matrix(v0, v1, scale=True) (Gohlke's
    transformations.py, which the reference asks the user to drop into hulo_transform/ and does
    not vendor -- PARITY UNPINNED): similarity transform (4 x 4) taking the 3 x n points v0 onto v1,
    rotation from the SVD of the covariance (Kabsch, reflection fixed), scale = sqrt of the ratio of
    the centred sums of squares."""
    v0 = np.array(v0, np.float64, copy=True)[:3]
    v1 = np.array(v1, np.float64, copy=True)[:3]
    t0 = -np.mean(v0, axis=1)
    t1 = -np.mean(v1, axis=1)
    v0 += t0.reshape(3, 1)
    v1 += t1.reshape(3, 1)
    u, s, vh = np.linalg.svd(np.dot(v1, v0.T))
    R = np.dot(u, vh)
    if np.linalg.det(R) < 0.0:
        R -= np.outer(u[:, 2], vh[2, :] * 2.0)
    M = np.identity(4)
    M[:3, :3] = R * np.sqrt(np.sum(v1 * v1) / np.sum(v0 * v0))
    M0 = np.identity(4); M0[:3, 3] = t0
    M1 = np.identity(4); M1[:3, 3] = t1
    return np.dot(np.linalg.inv(M1), np.dot(M, M0))


def ransac_transform3d(A, B, thres, samples, svd_ratio=float("inf"), similarity=False):
    """ransacAffineTransform (PyVisionLocalizeCommon/src/hulo_sfm/mergeSfM.py:344-388) and
    ransacSimilarityTransform (hulo_transform/ransacTransform.py:13-49) with the 4-point samples
    of every round given explicitly (the reference draws them with random.sample): find the 3 x 4
    M with A ~ M [B; 1].  Returns (M, inliers) or (empty, empty).  The affine flavour is pinned
    against the reference's own function (tests/golden/make_golden_merge.py)."""
    A = np.asarray(A, np.float64); B = np.asarray(B, np.float64)
    Bh = np.vstack((B, np.ones((1, B.shape[1]))))
    inliers = np.asarray([], np.int64)
    n_inl = 0
    for sel in np.asarray(samples, np.int64).reshape(-1, 4):
        if similarity:
            M = superimposition_matrix(B[:, sel], A[:, sel])[:3]
        else:
            M = np.linalg.lstsq(Bh[:, sel].T, A[:, sel].T, rcond=-1)[0].T
        norm = np.linalg.norm(np.dot(M, Bh) - A, axis=0)
        tmp = np.where(norm < thres)[0]
        if len(tmp) >= n_inl:
            s = np.linalg.svd(M[0:3, 0:3], compute_uv=False)
            if len(tmp) > n_inl and s[0] / s[-1] < svd_ratio:
                n_inl = len(tmp)
                inliers = tmp
    if len(inliers) < 4:
        return np.array([]), np.asarray([], np.int64)
    if similarity:
        M = superimposition_matrix(B[:, inliers], A[:, inliers])[:3]
    else:
        M = np.linalg.lstsq(Bh[:, inliers].T, A[:, inliers].T, rcond=-1)[0].T
    return M, inliers


# ---------------------------------------------------------------- BoF view selection
def bow_knn(bof, query, knn, subset=None):
    """Exact restatement of hulo::selectViewByBoF (BoWCommon/src/BoFUtils.cpp:27-68): the knn rows of
    `bof` (n x d float32, one per view) nearest to `query` under squared L2, by (distance, index).
    The reference searches a FLANN KD-tree (4 trees, 64 checks; OpenCV 3.0, not vendored), which is
    approximate; this is the exact answer it approximates, pinned against cv2.BFMatcher(NORM_L2)
    (tests/golden/make_golden_bow.py).  float32 accumulation in the order of the device kernel is
    NOT reproduced: distances are float64 sums rounded to float32, ties at that rounding are by index."""
    bof = np.asarray(bof, np.float32); query = np.asarray(query, np.float32).ravel()
    idx = np.arange(bof.shape[0]) if subset is None else np.asarray(subset, np.int64)
    d = ((bof[idx].astype(np.float64) - query.astype(np.float64)) ** 2).sum(axis=1)
    order = np.lexsort((idx, d))[:knn]
    return idx[order].astype(np.int32), d[order].astype(np.float32)
