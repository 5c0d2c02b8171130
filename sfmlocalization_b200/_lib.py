"""ctypes binding of libhulo_gpu.so (include/hulo_gpu.h).  No PyTorch, no CPU fallback:
importing works anywhere (so the symbol table can be checked on a CPU box) but every
compute entry point needs a B200 and raises HuloError otherwise."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhulo_gpu.so")

OK, ERR_ARG, ERR_CUDA, ERR_NCCL, ERR_CAPACITY = 0, 1, 2, 3, 4
PAIR_ONE_TO_ONE, PAIR_DROP_LAST = 1, 2
PAIR_REFERENCE = PAIR_ONE_TO_ONE | PAIR_DROP_LAST
KNN_INT, KNN_TC, KNN_AUTO, KNN_TC8 = 0, 1, 2, 3
DIST_NONE = 2**31 - 1
IDX_NONE = -1


class HuloError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("hulo_gpu status %d: %s" % (status, message))
        self.status = status


_vp, _sz, _i32, _u32, _u64, _f32, _f64 = (C.c_void_p, C.c_size_t, C.c_int32, C.c_uint32, C.c_uint64,
                                          C.c_float, C.c_double)
_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every symbol include/hulo_gpu.h declares
SIGNATURES = {
    "hulo_device_count": (C.c_int, []),
    "hulo_gpu_create": (C.c_int, [C.c_int, _pp]),
    "hulo_gpu_destroy": (None, [_vp]),
    "hulo_last_error": (C.c_char_p, []),
    "hulo_version": (C.c_char_p, []),
    "hulo_host_alloc": (C.c_int, [_sz, _pp]),
    "hulo_host_free": (None, [_vp]),
    "hulo_timer_start": (C.c_int, [_vp]),
    "hulo_timer_stop": (C.c_int, [_vp, C.POINTER(_f32)]),
    "hulo_synchronize": (C.c_int, [_vp]),
    "hulo_launch_count": (_u64, [_vp]),
    "hulo_gpu_set_knn_engine": (C.c_int, [_vp, C.c_int]),
    "hulo_gpu_knn_engine": (C.c_int, [_vp]),
    "hulo_db_upload": (C.c_int, [_vp, _vp, _sz, _sz, _vp, _sz, _pp]),
    "hulo_db_update": (C.c_int, [_vp, _vp, _vp, _sz, _sz]),
    "hulo_db_free": (None, [_vp]),
    "hulo_db_rows": (_sz, [_vp]),
    "hulo_db_segments": (_sz, [_vp]),
    "hulo_db_download": (C.c_int, [_vp, _vp, _sz, _sz, _vp]),
    "hulo_knn2": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "hulo_knn2_fetch": (C.c_int, [_vp, _sz, _vp, _vp]),
    "hulo_knn2_host": (C.c_int, [_vp, _vp, _sz, _sz, _vp, _sz, _sz, _vp, _vp]),
    "hulo_match_to_query": (C.c_int, [_vp, _vp, _vp, _sz, _vp, _sz, _sz, _f32, _vp, _vp, _vp, _vp, _sz,
                                      C.POINTER(_sz), _vp]),
    "hulo_match_to_queries": (C.c_int, [_vp, _vp, _vp, _sz, _vp, _sz, _vp, _sz, _f32, _vp, _vp, _vp, _vp, _vp, _sz,
                                        C.POINTER(_sz), _vp]),
    "hulo_match_pairs": (C.c_int, [_vp, _vp, _vp, _sz, _f32, C.c_uint, _vp, _vp, _vp, _sz, C.POINTER(_sz)]),
    "hulo_score_resection": (C.c_int, [_vp, _vp, _sz, _vp, _vp, _sz, _vp, _f64, _vp, _vp, _vp, _vp]),
    "hulo_resection_residuals": (C.c_int, [_vp, _vp, _sz, _vp, _vp, _sz, _vp, _vp]),
    "hulo_p3p": (C.c_int, [_vp, _vp, _sz, _vp, _vp, _sz, _vp, _vp, _vp]),
    "hulo_resect_acransac": (C.c_int, [_vp, _vp, _vp, _sz, _vp, _sz, _u64, _vp, _vp, C.POINTER(_sz),
                                       C.POINTER(_f64), C.POINTER(C.c_int)]),
    "hulo_resect_acransac_sequential": (C.c_int, [_vp, _vp, _vp, _sz, _vp, _sz, _u64, _vp, _vp, C.POINTER(_sz),
                                                  C.POINTER(_f64), C.POINTER(C.c_int)]),
    "hulo_selftest_rescoring": (C.c_int, [_u64, _sz, _sz, C.POINTER(_sz)]),
    "hulo_resect_acransac_batch": (C.c_int, [_vp, _sz, _vp, _vp, _vp, _vp, _sz, _u64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hulo_pose_from_projection": (C.c_int, [_vp, _vp, _vp, _vp]),
    "hulo_geometric_filter": (C.c_int, [_vp, _vp, _vp, _vp, _sz, _vp, _f64, _sz, _u64, _vp, _vp, _vp, _vp, _vp, _vp,
                                        _vp]),
    "hulo_guided_match": (C.c_int, [_vp, _vp, _vp, _vp, _sz, _vp, _vp, _f64, C.c_int, _vp, _vp, _vp, _sz,
                                    C.POINTER(_sz)]),
    "hulo_ransac_transform3d": (C.c_int, [_vp, _vp, _vp, _sz, _f64, _vp, _sz, _u64, _f64, C.c_int, _vp, _vp,
                                          C.POINTER(_sz), _vp]),
    "hulo_bow_create": (C.c_int, [_vp, _vp, _sz, _sz, _pp]),
    "hulo_bow_destroy": (None, [_vp]),
    "hulo_bow_knn": (C.c_int, [_vp, _vp, _vp, _sz, _sz, _vp, _vp]),
    "hulo_engine_create": (C.c_int, [_vp, _vp, _sz, _sz, _vp, _sz, _vp, _vp, _vp, _sz, _vp, _sz, _vp, _pp]),
    "hulo_engine_destroy": (None, [_vp]),
    "hulo_engine_configure": (C.c_int, [_vp, _f32, C.c_int, C.c_int, C.c_int, _sz]),
    "hulo_engine_set_resection_schedule": (C.c_int, [_vp, C.c_int]),
    "hulo_engine_set_keypoints": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int]),
    "hulo_engine_set_query_size": (C.c_int, [_vp, C.c_int, C.c_int]),
    "hulo_engine_configure_geometric": (C.c_int, [_vp, C.c_int, _sz, _f64]),
    "hulo_engine_set_guided_matching": (C.c_int, [_vp, C.c_int]),
    "hulo_engine_localize": (C.c_int, [_vp, _vp, _sz, _sz, _vp, _vp, _sz, _u64, _vp, C.POINTER(C.c_int), _vp, _vp,
                                       C.POINTER(_sz), _vp, C.POINTER(_sz), _vp]),
    "hulo_engine_localize_sharded": (C.c_int, [_vp, _vp, _sz, _sz, _vp, _vp, _sz, _u64, _vp, C.POINTER(C.c_int), _vp, _vp,
                                               C.POINTER(_sz), _vp, C.POINTER(_sz), _vp]),
    "hulo_partition_views": (C.c_int, [_vp, _sz, C.c_int, _vp]),
    "hulo_comm_allgather": (C.c_int, [_vp, _vp, _sz, _vp]),
    "hulo_comm_exchange_kind": (C.c_char_p, [_vp]),
    "hulo_comm_rank": (C.c_int, [_vp]),
    "hulo_comm_world": (C.c_int, [_vp]),
    "hulo_engine_localize_batch": (C.c_int, [_vp, _sz, _vp, _sz, _vp, _vp, _vp, _sz, _u64, _vp, _vp, _vp, _vp, _vp]),
    "hulo_comm_unique_id": (C.c_int, [_vp]),
    "hulo_comm_init": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "hulo_comm_barrier": (C.c_int, [_vp]),
    "hulo_comm_max_f64": (C.c_int, [_vp, C.POINTER(_f64)]),
    "hulo_knn2_sharded": (C.c_int, [_vp, _vp, _vp, _u64, _vp, _vp]),
    "hulo_knn2_sharded_submit": (C.c_int, [_vp, _vp, _vp, _u64]),
    "hulo_knn2_sharded_collect": (C.c_int, [_vp, _vp, _vp, C.POINTER(_sz)]),
    "hulo_merge_top2": (C.c_int, [_vp, _vp, _sz, C.c_int, _vp, _vp]),
}

_lib = None


def load():
    """dlopen the in-tree library; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HuloError(ERR_CUDA, "%s is missing: build it with `make -C sfmlocalization_b200/csrc` "
                                      "(or __graft_entry__.build()); there is no fallback path" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)      # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status):
    if status != OK:
        raise HuloError(status, load().hulo_last_error().decode("utf-8", "replace"))
