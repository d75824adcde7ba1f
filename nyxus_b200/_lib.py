"""ctypes binding of include/nyxus_b200.h (libnyxus_b200.so, built in-tree by nyxus_b200/build.py).

The library is the product; there is no Python/CPU fallback: if it is missing or no sm_100 device is visible,
every entry point raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnyxus_b200.so")

NSTAGES = 18


class Settings(C.Structure):
    _fields_ = [("device", C.c_int32), ("coarse_gray_depth", C.c_int32), ("ibsi", C.c_int32), ("glcm_offset", C.c_int32),
                ("glcm_symmetric", C.c_int32), ("verbose", C.c_int32), ("soft_nan", C.c_double), ("dim", C.c_int32),
                ("reserved", C.c_int32)]


class Table(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_double)), ("nrows", C.c_int64), ("ncols", C.c_int32), ("reserved", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("nyxus_b200: %s is missing -- build it with `python -m nyxus_b200.build` "
                           "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    L.nyxb200_default_settings.argtypes = [C.POINTER(Settings)]
    L.nyxb200_create.argtypes = [C.POINTER(Settings), C.POINTER(vp)]
    L.nyxb200_destroy.argtypes = [vp]
    L.nyxb200_last_error.restype = C.c_char_p
    L.nyxb200_last_error.argtypes = [vp]
    L.nyxb200_select_features.argtypes = [vp, C.POINTER(C.c_char_p), i32]
    L.nyxb200_num_columns.argtypes = [vp]
    L.nyxb200_column_name.restype = C.c_char_p
    L.nyxb200_column_name.argtypes = [vp, i32]
    L.nyxb200_column_code.argtypes = [vp, i32]
    L.nyxb200_set_param.argtypes = [vp, C.c_char_p, C.c_double]
    for f in (L.nyxb200_featurize_2d_u32, L.nyxb200_featurize_2d_u16):
        f.argtypes = [vp, vp, vp, i32, i64, i64, C.POINTER(Table), C.POINTER(C.POINTER(C.c_int32))]
    L.nyxb200_featurize_2d_device.argtypes = [vp, vp, i32, vp, i32, i64, i64, C.POINTER(Table), C.POINTER(C.POINTER(C.c_int32))]
    for f in (L.nyxb200_featurize_3d_u32, L.nyxb200_featurize_3d_u16):
        f.argtypes = [vp, vp, vp, i32, i64, i64, i64, C.POINTER(Table), C.POINTER(C.POINTER(C.c_int32))]
    L.nyxb200_featurize_3d_device.argtypes = [vp, vp, i32, vp, i32, i64, i64, i64, C.POINTER(Table), C.POINTER(C.POINTER(C.c_int32))]
    L.nyxb200_featurize_2d_list.argtypes = [vp, C.POINTER(vp), i32, C.POINTER(vp), i32, i32, i64, i64, C.POINTER(Table), C.POINTER(C.POINTER(C.c_int32))]
    L.nyxb200_featurize_3d_list.argtypes = [vp, C.POINTER(vp), i32, C.POINTER(vp), i32, i32, i64, i64, i64, C.POINTER(Table), C.POINTER(C.POINTER(C.c_int32))]
    L.nyxb200_table_free.argtypes = [vp, C.POINTER(Table)]
    L.nyxb200_table_detach.argtypes = [vp, C.POINTER(Table)]
    L.nyxb200_host_release.argtypes = [vp]
    L.nyxb200_host_release.restype = None
    L.nyxb200_validate_features.argtypes = [i32, C.POINTER(C.c_char_p), i32, C.c_char_p, C.c_size_t]
    L.nyxb200_get_param.argtypes = [vp, C.c_char_p, C.POINTER(C.c_double)]
    L.nyxb200_device_count.argtypes = []
    L.nyxb200_device_info.argtypes = [i32, C.c_char_p, C.c_size_t, C.POINTER(i64), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.nyxb200_shard_slides.argtypes = [i32, i32, C.POINTER(i64), C.POINTER(C.c_int32)]
    L.nyxb200_featurize_2d_list_multi.argtypes = [C.POINTER(vp), i32, C.POINTER(vp), i32, C.POINTER(vp), i32, i32, i64, i64,
                                                  C.POINTER(i64), C.POINTER(Table), C.POINTER(C.POINTER(C.c_int32))]
    L.nyxb200_last_timings.argtypes = [vp, C.POINTER(C.c_float), i32]
    L.nyxb200_last_counts.argtypes = [vp, C.POINTER(C.c_int64), i32]
    L.nyxb200_stage_name.restype = C.c_char_p
    L.nyxb200_stage_name.argtypes = [i32]
    L.nyxb200_debug_rois.argtypes = [vp, vp, i64, C.POINTER(C.c_int64)]
    L.nyxb200_debug_cloud.argtypes = [vp, C.c_uint32, vp, i64, C.POINTER(C.c_int64)]
    L.nyxb200_debug_matrix.argtypes = [vp, C.c_uint32, i32, vp, i64, C.POINTER(C.c_int64), vp]
    L.nyxb200_family_id.argtypes = [C.c_char_p]
    L.nyxb200_version.restype = C.c_char_p
    _lib = L
    return L


EXPORTS = ["nyxb200_default_settings", "nyxb200_create", "nyxb200_destroy", "nyxb200_last_error", "nyxb200_select_features",
           "nyxb200_num_columns", "nyxb200_column_name", "nyxb200_column_code", "nyxb200_set_param", "nyxb200_featurize_2d_u32",
           "nyxb200_featurize_2d_u16", "nyxb200_featurize_2d_device", "nyxb200_featurize_3d_u32", "nyxb200_featurize_3d_u16",
           "nyxb200_featurize_3d_device", "nyxb200_featurize_2d_list", "nyxb200_featurize_3d_list", "nyxb200_table_free", "nyxb200_last_timings",
           "nyxb200_last_counts", "nyxb200_stage_name", "nyxb200_debug_rois", "nyxb200_debug_cloud", "nyxb200_debug_matrix",
           "nyxb200_family_id", "nyxb200_version", "nyxb200_table_detach", "nyxb200_host_release", "nyxb200_validate_features",
           "nyxb200_get_param", "nyxb200_device_count", "nyxb200_device_info", "nyxb200_shard_slides",
           "nyxb200_featurize_2d_list_multi"]
