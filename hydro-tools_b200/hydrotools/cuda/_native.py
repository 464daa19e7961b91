"""ctypes binding of libhydrotools_b200.so (C-ABI: include/hydrotools_b200.h).

There is NO CPU fallback: if the shared library is missing this module raises
at import, and every compute entry point raises if CUDA is unavailable.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libhydrotools_b200.so")

HT_TERRAIN_NODATA = -9999.0

DIR_CODE = 0x0F
DIR_NEG = 0x10
DIR_EDGE = 0x20
DIR_TAINT = 0x40
DIR_NULL = 0x80


class NativeError(RuntimeError):
    """Non-zero return from the C-ABI."""


if not os.path.isfile(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python hydro-tools_b200/build.py` "
        "(nvcc, sm_100a). hydrotools.cuda has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_p = C.c_void_p
_i64 = C.c_int64
_int = C.c_int
_f32 = C.c_float
_f64 = C.c_double
_u64 = C.c_uint64

lib.ht_version.restype = _int
lib.ht_error_string.restype = C.c_char_p
lib.ht_error_string.argtypes = [_int]

SIGNATURES = {
    "ht_terrain_fused_f32": [_p, _i64, _i64, _i64, _int, _f32, _f64, _f64, _f64, _int,
                             _p, _p, _p, _i64, _int, _int, _p],
    "ht_window_filter_f32": [_p, _i64, _i64, _i64, _int, _f32, _p, _int, _int, _int, _p, _i64,
                             _int, _int, _p],
    "ht_window_filter_runs_f32": [_p, _i64, _i64, _i64, _int, _f32, _p, _int, _int, _int, _p,
                                  _i64, _int, _int, _p, C.c_size_t, _p],
    "ht_d8_f32": [_p, _i64, _i64, _i64, _int, _f32, _f64, _f64, _i64, _i64, _int, _int,
                  _p, _i64, _p],
    "ht_d8_part_f32": [_p, _i64, _i64, _i64, _int, _f32, _f64, _f64, _i64, _i64, _int, _int,
                       _int, _int, _p, _i64, _p],
    "ht_d8_validate_conditioned_f32": [_p, _i64, _i64, _i64, _int, _f32, _i64, _i64, _int,
                                       _int, _p, _p],
    "ht_d8_general_f32": [_p, _i64, _i64, _i64, _int, _f32, _f64, _f64, _p, _i64, _p, C.c_size_t,
                          C.POINTER(C.c_longlong), _p],
    "ht_dir_pack_i8": [_p, _i64, _i64, _i64, _p, _i64, _p],
    "ht_dir_pack_band_i8": [_p, _i64, _i64, _i64, _i64, _i64, _int, _int, _p, _i64, _p],
    "ht_dir_unpack_i8": [_p, _i64, _i64, _i64, _p, _i64, _p],
    "ht_acc_local": [_p, _i64, _i64, _i64, _i64, _i64, _int, _int, _int, _p, _i64, _int,
                     _f32, _p, C.c_size_t, _p],
    "ht_acc_resolve": [_i64, _i64, _int, _int, _p, C.c_size_t, _p, _p],
    "ht_acc_resolve_delta": [_i64, _i64, _int, _p, _p, _p, C.c_size_t, _p],
    "ht_acc_mark_edges": [_i64, _i64, _int, _int, _int, _p, C.c_size_t, _p],
    "ht_acc_add_imports": [_i64, _i64, _int, _p, _p, _p, C.c_size_t, _p],
    "ht_acc_weight_max": [_p, _i64, _i64, _i64, _int, _f32, _p, _p],
    "ht_acc_set_weight_max": [_i64, _i64, _p, C.c_size_t, _p, _p],
    "ht_band_tables": [_i64, _i64, _int, _p, C.c_size_t, _p, _p],
    "ht_band_graph": [_p, _int, _i64, _int, _p, _p, _p],
    "ht_acc_final": [_p, _i64, _i64, _i64, _i64, _i64, _int, _int, _int, _p, _i64, _int,
                     _f32, _p, _i64, _p, _i64, _p, C.c_size_t, _p],
    "ht_acc_stage_b_mode": [_int],
    "ht_acc_tile_phases": [_i64, _i64, _int, _p, C.c_size_t, _p],
    "ht_acc_final_phase": [_p, _i64, _i64, _i64, _i64, _i64, _int, _int, _int, _p, _i64, _int,
                           _f32, _p, _i64, _p, _i64, _p, C.c_size_t, _int, _p],
    "ht_acc_f64": [_p, _i64, _i64, _i64, _int, _p, _i64, _int, _f32, _p, _i64, _p, _i64,
                   _p, C.c_size_t, _p],
    "ht_acc_workspace_layout": [_i64, _i64, _p, C.c_size_t, C.POINTER(C.c_int64)],
    "ht_forest_resolve": [_p, _p, _i64, _int, _p, _p, C.c_size_t, _p],
    "ht_acc_apply_sign": [_p, _i64, _p, _i64, _p, _i64, _i64, _i64, _p],
    "ht_reach_labels_i32": [_p, _i64, _p, _i64, _i64, _i64, C.c_int32, _p, _i64, _p, C.c_size_t, _p],
    "ht_upstream_mean_f64": [_p, _i64, _p, _i64, _p, _i64, _i64, _i64, _f64, _f64, _f64, _int, _i64,
                             _p, _i64, _p, C.c_size_t, C.POINTER(C.c_longlong), _p],
    "ht_basin_u8": [_p, _i64, _i64, _i64, _i64, _i64, _p, _i64, _p, C.c_size_t, _p],
    "ht_tiff_encode_tiles": [_p, _i64, _i64, _i64, _int, _u64, _p, C.c_size_t, _p, _p, _p, _p,
                             C.c_size_t, _p],
    "ht_synth_dem_f32": [_p, _i64, _i64, _i64, _u64, _i64, _i64, _int, _f32, _p],
    "ht_synth_weight_f32": [_p, _i64, _i64, _i64, _u64, _i64, _p],
}


def _declare():
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _int


_declare()
lib.ht_acc_workspace_bytes.argtypes = [_i64, _i64]
lib.ht_acc_workspace_bytes.restype = C.c_size_t
lib.ht_acc_workspace_bytes_mode.argtypes = [_i64, _i64, _int]
lib.ht_acc_workspace_bytes_mode.restype = C.c_size_t
lib.ht_tiff_encode_bound.argtypes = [_i64, _i64, _int]
lib.ht_tiff_encode_bound.restype = C.c_size_t
lib.ht_tiff_encode_workspace_bytes.argtypes = [_i64, _i64, _int]
lib.ht_tiff_encode_workspace_bytes.restype = C.c_size_t
lib.ht_streams_workspace_bytes.argtypes = [_i64, _i64, _int, _i64]
lib.ht_streams_workspace_bytes.restype = C.c_size_t
lib.ht_d8_general_workspace_bytes.argtypes = [_i64, _i64]
lib.ht_d8_general_workspace_bytes.restype = C.c_size_t
lib.ht_window_filter_runs_workspace_bytes.argtypes = [_i64, _i64, _int]
lib.ht_window_filter_runs_workspace_bytes.restype = C.c_size_t

MODE_COUNT, MODE_TAINT, MODE_WEIGHTED = 0, 1, 2


def check(rc: int, what: str):
    if rc == 0:
        return
    msg = lib.ht_error_string(rc).decode()
    if rc in (-1, -2):
        raise ValueError(f"{what}: {msg} (code {rc})")
    raise NativeError(f"{what}: {msg} (code {rc})")


def call(name: str, *args):
    check(getattr(lib, name)(*args), name)
