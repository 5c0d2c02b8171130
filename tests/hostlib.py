"""ctypes access to libhulo_host.so's test surface (csrc/host/host_capi.cpp)."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "sfmlocalization_b200")
CLI = os.path.join(PKG, "hulo_ext_match")
CLI_LOCALIZE = os.path.join(PKG, "hulo_localize")
CLI_BA_RESECT = os.path.join(PKG, "hulo_ba_resect")
_lib = None


def lib():
    global _lib
    if _lib is None:
        C.CDLL(os.path.join(PKG, "libhulo_gpu.so"), mode=C.RTLD_GLOBAL)
        _lib = C.CDLL(os.path.join(PKG, "libhulo_host.so"))
        for name in ("hulo_host_read_desc", "hulo_host_matches_roundtrip", "hulo_host_views_from_sfm_data",
                     "hulo_host_load_sfm_data", "hulo_host_read_feat", "hulo_host_read_mat_bin",
                     "hulo_host_select_view_by_bof"):
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


def write_feat(path, xy):
    """OpenMVG .feat (SIOPointFeature text): x y scale orientation per line."""
    with open(path, "w") as f:
        for x, y in np.asarray(xy, np.float64):
            f.write("%r %r 1.0 0.0\n" % (float(x), float(y)))


def write_sfm_data(path, scene, names, focal=None, disto=None):
    """An OpenMVG 1.x cereal sfm_data.json for a synth.localization_scene: views with sizes,
    one pinhole (or pinhole_radial_k3) intrinsic, one pose per view, landmarks with observations."""
    import json
    K = scene["K"]
    V = len(scene["seg_offsets"]) - 1
    w, h = [int(x) for x in scene["view_wh"][0]]
    views = [{"key": k, "value": {"polymorphic_id": 1073741824, "ptr_wrapper": {"id": 2147483649 + k, "data": {
        "local_path": "/", "filename": names[k] + ".jpg", "width": w, "height": h, "id_view": k, "id_intrinsic": 0,
        "id_pose": k}}}} for k in range(V)]
    data = {"width": w, "height": h, "focal_length": float(K[0, 0] if focal is None else focal),
            "principal_point": [float(K[0, 2]), float(K[1, 2])]}
    name = "pinhole"
    if disto is not None:
        data["disto_k3"] = [float(x) for x in disto]
        name = "pinhole_radial_k3"
    intr = [{"key": 0, "value": {"polymorphic_id": 2147483649, "polymorphic_name": name,
                                 "ptr_wrapper": {"id": 2147483700, "data": data}}}]
    ext = [{"key": k, "value": {"rotation": scene["view_R"][k].tolist(),
                                "center": (-scene["view_R"][k].T @ scene["view_t"][k]).tolist()}} for k in range(V)]
    obs = {}
    off = scene["seg_offsets"].astype(np.int64)
    for v, f, l in zip(scene["obs_view"].tolist(), scene["obs_feat"].tolist(), scene["obs_landmark"].tolist()):
        obs.setdefault(l, []).append({"key": v, "value": {"id_feat": f, "x": scene["map_xy"][off[v] + f].tolist()}})
    structure = [{"key": l, "value": {"X": scene["landmark_X"][l].tolist(), "observations": obs[l]}}
                 for l in sorted(obs)]
    with open(path, "w") as f:
        json.dump({"sfm_data_version": "0.2", "root_path": "/x", "views": views, "intrinsics": intr,
                   "extrinsics": ext, "structure": structure, "control_points": []}, f)
    return sorted(obs)


def load_sfm_data(path, cap_views=4096):
    counts = np.zeros(5, np.uint64); X = np.zeros(3); intr = np.zeros(8)
    wh = np.zeros((cap_views, 2), np.uint64)
    n = lib().hulo_host_load_sfm_data(path.encode(), _p(counts), _p(X), _p(intr), _p(wh), C.c_ulonglong(cap_views))
    if n < 0:
        return None
    return dict(counts=counts.astype(np.int64), first_X=X, intrinsic0=intr, view_wh=wh[:n].astype(np.int64))


def sfm_observation_sums(path):
    out = np.zeros(2)
    return out if lib().hulo_host_sfm_observation_sums(path.encode(), _p(out)) == 0 else None


def save_sfm_poses(in_json, out_json, ids, R, center):
    ids = np.ascontiguousarray(ids, np.uint64); R = np.ascontiguousarray(R, np.float64)
    center = np.ascontiguousarray(center, np.float64)
    return lib().hulo_host_save_sfm_poses(in_json.encode(), out_json.encode(), C.c_ulonglong(len(ids)), _p(ids), _p(R),
                                          _p(center)) == 0


def undistort(focal, ppx, ppy, k, x, y):
    out = np.zeros(2)
    lib().hulo_host_undistort(C.c_double(focal), C.c_double(ppx), C.c_double(ppy), C.c_double(k[0]), C.c_double(k[1]),
                              C.c_double(k[2]), C.c_double(x), C.c_double(y), _p(out))
    return out


def read_cv_matrix(path, name, cap=64):
    out = np.zeros(cap); r = C.c_int(0); c = C.c_int(0)
    if lib().hulo_host_read_cv_matrix(path.encode(), name.encode(), _p(out), cap, C.byref(r), C.byref(c)) != 0:
        return None
    return out[:r.value * c.value].reshape(r.value, c.value)


def read_feat(path):
    n = lib().hulo_host_read_feat(path.encode(), None, C.c_ulonglong(0))
    if n < 0:
        return None
    out = np.zeros((max(n, 1), 2))
    lib().hulo_host_read_feat(path.encode(), _p(out), C.c_ulonglong(n))
    return out[:n]


def save_mat_bin(path, mat, cv_type):
    """hulo::saveMatBin (FileUtils.cpp:44-58): cv_type 0 = CV_8U, 4 = CV_32S, 5 = CV_32F, 6 = CV_64F."""
    m = np.ascontiguousarray(mat, np.float64)
    rows, cols = (m.shape if m.ndim == 2 else (m.shape[0], 1)) if m.size else (0, 0)
    return lib().hulo_host_save_mat_bin(path.encode(), rows, cols, cv_type, _p(m))


def read_mat_bin(path):
    r = C.c_int(0); c = C.c_int(0)
    n = lib().hulo_host_read_mat_bin(path.encode(), C.byref(r), C.byref(c), None, C.c_ulonglong(0))
    if n < 0:
        return None
    out = np.zeros(max(n, 1))
    lib().hulo_host_read_mat_bin(path.encode(), C.byref(r), C.byref(c), _p(out), C.c_ulonglong(n))
    return out[:n].reshape(r.value, c.value)


def select_view_by_bof(match_dir, n_views, bow, view_list, knn):
    bow = np.ascontiguousarray(bow, np.float32)
    vl = np.ascontiguousarray(view_list, np.uint64)
    out = np.zeros(max(knn, 1), np.uint64)
    n = lib().hulo_host_select_view_by_bof(match_dir.encode(), C.c_ulonglong(n_views), _p(bow), C.c_ulonglong(len(bow)),
                                           _p(vl), C.c_ulonglong(len(vl)), knn, _p(out))
    return None if n < 0 else out[:n].astype(np.int64)
