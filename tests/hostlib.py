"""ctypes access to libhulo_host.so's test surface (csrc/host/host_capi.cpp)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sfmlocalization_b200")
CLI = os.path.join(PKG, "hulo_ext_match")
_lib = None


def lib():
    global _lib
    if _lib is None:
        C.CDLL(os.path.join(PKG, "libhulo_gpu.so"), mode=C.RTLD_GLOBAL)
        _lib = C.CDLL(os.path.join(PKG, "libhulo_host.so"))
        for name in ("hulo_host_read_desc", "hulo_host_matches_roundtrip", "hulo_host_views_from_sfm_data"):
            getattr(_lib, name).restype = C.c_longlong
        for name in ("hulo_host_all_pairs", "hulo_host_video_pairs", "hulo_host_remove_dup_pairs",
                     "hulo_host_partition_pairs", "hulo_host_propagate_tracks"):
            getattr(_lib, name).restype = C.c_ulonglong
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def write_desc_numpy(path, rows):
    """The .desc layout: uint64 LE count + count x 64 bytes."""
    rows = np.ascontiguousarray(rows, np.uint8)
    r64 = np.zeros((rows.shape[0], 64), np.uint8)
    r64[:, :rows.shape[1]] = rows
    with open(path, "wb") as f:
        f.write(np.uint64(rows.shape[0]).tobytes())
        f.write(r64.tobytes())


def read_desc(path):
    n = lib().hulo_host_read_desc(path.encode(), None, C.c_ulonglong(0))
    if n < 0:
        return None
    out = np.empty((n, 64), np.uint8)
    lib().hulo_host_read_desc(path.encode(), _p(out), C.c_ulonglong(n))
    return out


def write_desc(path, rows):
    rows = np.ascontiguousarray(rows, np.uint8)
    return lib().hulo_host_write_desc(path.encode(), _p(rows), C.c_ulonglong(rows.shape[0]),
                                      C.c_ulonglong(rows.shape[1]))


def all_pairs(ids):
    ids = np.ascontiguousarray(ids, np.uint64)
    cap = len(ids) * len(ids) + 1
    out = np.empty((cap, 2), np.uint64)
    n = lib().hulo_host_all_pairs(_p(ids), C.c_ulonglong(len(ids)), _p(out), C.c_ulonglong(cap))
    return out[:n].astype(np.int64)


def video_pairs(ids, frame):
    ids = np.ascontiguousarray(ids, np.uint64)
    cap = len(ids) * len(ids) + 1
    out = np.empty((cap, 2), np.uint64)
    n = lib().hulo_host_video_pairs(_p(ids), C.c_ulonglong(len(ids)), frame, _p(out), C.c_ulonglong(cap))
    return out[:n].astype(np.int64)


def remove_dup_pairs(pairs):
    p = np.ascontiguousarray(pairs, np.uint64).reshape(-1, 2).copy()
    n = lib().hulo_host_remove_dup_pairs(_p(p), C.c_ulonglong(len(p)))
    return p[:n].astype(np.int64)


def partition_pairs(pairs, rows, rank, world):
    p = np.ascontiguousarray(pairs, np.uint64).reshape(-1, 2)
    rows = np.ascontiguousarray(rows, np.uint64)
    out = np.empty(max(len(p), 1), np.uint64)
    n = lib().hulo_host_partition_pairs(_p(p), C.c_ulonglong(len(p)), _p(rows), C.c_ulonglong(len(rows)), rank,
                                        world, _p(out))
    return out[:n].astype(np.int64)


def propagate_tracks(n_frames, max_dist, feat_number, m_off, m_i, m_j):
    fn = np.ascontiguousarray(feat_number, np.int32); mo = np.ascontiguousarray(m_off, np.int64)
    mi = np.ascontiguousarray(m_i, np.int32); mj = np.ascontiguousarray(m_j, np.int32)
    n = lib().hulo_host_propagate_tracks(C.c_ulonglong(n_frames), C.c_ulonglong(max_dist), _p(fn), _p(mo), _p(mi),
                                         _p(mj), None, C.c_ulonglong(0))
    out = np.empty((max(n, 1), 4), np.int32)
    lib().hulo_host_propagate_tracks(C.c_ulonglong(n_frames), C.c_ulonglong(max_dist), _p(fn), _p(mo), _p(mi),
                                     _p(mj), _p(out), C.c_ulonglong(n))
    return out[:n]


def parse_matches(path):
    """OpenMVG text match file -> {(I, J): [(i, j), ...]} in file order."""
    out = {}
    with open(path) as f:
        tok = f.read().split()
    k = 0
    while k < len(tok):
        I, J, n = int(tok[k]), int(tok[k + 1]), int(tok[k + 2])
        k += 3
        out[(I, J)] = [(int(tok[k + 2 * m]), int(tok[k + 2 * m + 1])) for m in range(n)]
        k += 2 * n
    return out
