"""ctypes binding of include/shapes_b200.h.

Loading fails loudly when the CUDA library has not been built: there is no
CPU or PyTorch fallback for this path.
"""
from __future__ import annotations

import ctypes as C
import os

_ROOT = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SHAPES_B200_LIB") or os.path.join(_ROOT, "lib", "libshapes_b200.so")

OK, E_ARG, E_CUDA, E_NCCL, E_CAPACITY = 0, -1, -2, -3, -4
NCCL_ID_BYTES = 128
IPC_BYTES = 2048
N_STAGES = 12

_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)
_f64p = C.POINTER(C.c_double)


class FrameOut(C.Structure):
    """shapes_frame_out"""
    _fields_ = [
        ("n_pairs", C.c_int64), ("pair_i", _i32p), ("pair_j", _i32p),
        ("n_contacts", C.c_int64),
        ("key_i", _i32p), ("key_j", _i32p), ("feat_a", _i32p), ("feat_b", _i32p),
        ("flip", _u8p),
        ("normal_x", _f64p), ("normal_y", _f64p), ("center_x", _f64p), ("center_y", _f64p),
        ("depth", _f64p),
        ("j_np", _f64p * 6), ("b_np", _f64p),
        ("ra_x", _f64p), ("ra_y", _f64p), ("rb_x", _f64p), ("rb_y", _f64p),
        ("rn_x", _f64p), ("rn_y", _f64p),
        ("j_f", _f64p * 6), ("b_f", _f64p),
        ("inv_eff_np", _f64p), ("inv_eff_f", _f64p),
        ("warm_np", _f64p), ("warm_f", _f64p), ("warm_hit", _u8p),
        ("aabb_min_x", _f64p), ("aabb_max_x", _f64p), ("aabb_min_y", _f64p), ("aabb_max_y", _f64p),
        ("world_x", _f64p), ("world_y", _f64p),
        ("n_big", C.c_int64), ("grid_w", C.c_int32), ("grid_h", C.c_int32),
        ("cell_size", C.c_double), ("device_ms", C.c_float), ("total_ms", C.c_float),
    ]


class DeviceView(C.Structure):
    """shapes_device_view"""
    _fields_ = [
        ("n_pairs", C.c_int64), ("n_contacts", C.c_int64),
        ("pair_i", C.c_void_p), ("pair_j", C.c_void_p),
        ("key_i", C.c_void_p), ("key_j", C.c_void_p), ("feat_a", C.c_void_p), ("feat_b", C.c_void_p),
        ("flip", C.c_void_p),
        ("normal_x", C.c_void_p), ("normal_y", C.c_void_p), ("center_x", C.c_void_p),
        ("center_y", C.c_void_p), ("depth", C.c_void_p),
        ("j_np", C.c_void_p * 6), ("b_np", C.c_void_p),
        ("ra_x", C.c_void_p), ("ra_y", C.c_void_p), ("rb_x", C.c_void_p), ("rb_y", C.c_void_p),
        ("rn_x", C.c_void_p), ("rn_y", C.c_void_p),
        ("j_f", C.c_void_p * 6),
        ("inv_eff_np", C.c_void_p), ("inv_eff_f", C.c_void_p),
        ("warm_np", C.c_void_p), ("warm_f", C.c_void_p), ("warm_hit", C.c_void_p),
        ("aabb", C.c_void_p),
    ]


class StepConfig(C.Structure):
    """shapes_step_config"""
    _fields_ = [
        ("dt", C.c_double), ("baumgarte", C.c_double), ("slop", C.c_double),
        ("external_kind", C.c_int32), ("solver_iterations", C.c_int32),
        ("external_x", C.c_double), ("external_y", C.c_double),
        ("warm_start", C.c_int32), ("reserved", C.c_int32),
    ]


class StepStats(C.Structure):
    """shapes_step_stats"""
    _fields_ = [
        ("n_pairs", C.c_int64), ("n_contacts", C.c_int64),
        ("solver_nodes", C.c_int64), ("queue_pushes", C.c_int64), ("body_chains", C.c_int64),
        ("warm", C.c_int32), ("reserved", C.c_int32),
        ("frame_ms", C.c_float), ("chains_ms", C.c_float), ("solve_ms", C.c_float),
        ("integrate_ms", C.c_float), ("total_ms", C.c_float),
    ]


class FrameInfo(C.Structure):
    """shapes_frame_info"""
    _fields_ = [("pairs_with_contacts", C.c_int64), ("sorted_mode", C.c_int32), ("sat_kernel", C.c_int32)]


SAT_KERNEL_NAMES = {0: "k_manifolds<4>", 1: "k_manifolds<8>", 2: "k_manifolds<8,circles>", 3: "k_manifolds_coop"}

EXT_NONE, EXT_ACCEL, EXT_FORCE = 0, 1, 2

# every symbol include/shapes_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "shapes_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
    "shapes_create_ranked": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p,
                                       C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
    "shapes_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "shapes_destroy": (None, [C.c_void_p]),
    "shapes_last_error": (C.c_char_p, [C.c_void_p]),
    "shapes_set_hulls": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "shapes_set_shapes": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "shapes_set_cell_size": (C.c_int, [C.c_void_p, C.c_double]),
    "shapes_grow": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64]),
    "shapes_frame": (C.c_int, [C.c_void_p, C.c_int64] + [C.c_void_p] * 7 + [C.c_double] * 3 + [C.POINTER(FrameOut)]),
    "shapes_frame_device": (C.c_int, [C.c_void_p, C.c_int64] + [C.c_void_p] * 7 + [C.c_double] * 3 + [C.POINTER(FrameOut)]),
    "shapes_device_view_get": (C.c_int, [C.c_void_p, C.POINTER(DeviceView)]),
    "shapes_fetch": (C.c_int, [C.c_void_p, C.POINTER(FrameOut)]),
    "shapes_set_lagrangian_cache": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "shapes_set_lagrangian_cache_device": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "shapes_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "shapes_ipc_import": (C.c_int, [C.c_void_p, C.c_void_p]),
    "shapes_rank_segments": (C.c_int, [C.c_void_p, C.c_int] + [C.POINTER(C.c_int64)] * 4),
    "shapes_create_multi": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
    "shapes_multi_destroy": (None, [C.c_void_p]),
    "shapes_multi_last_error": (C.c_char_p, [C.c_void_p]),
    "shapes_multi_set_shapes": (C.c_int, [C.c_void_p, C.c_int64] + [C.c_void_p] * 7),
    "shapes_multi_frame": (C.c_int, [C.c_void_p, C.c_int64] + [C.c_void_p] * 7 + [C.c_double] * 3 + [C.POINTER(FrameOut)]),
    "shapes_multi_rank": (C.c_void_p, [C.c_void_p, C.c_int]),
    "shapes_rank_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                   C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "shapes_world_upload": (C.c_int, [C.c_void_p, C.c_int64] + [C.c_void_p] * 12),
    "shapes_world_download": (C.c_int, [C.c_void_p, C.c_int64] + [C.c_void_p] * 8),
    "shapes_world_step": (C.c_int, [C.c_void_p, C.POINTER(StepConfig), C.POINTER(StepStats)]),
    "shapes_sincos": (None, [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "shapes_host_alloc": (C.c_void_p, [C.c_size_t]),
    "shapes_host_free": (None, [C.c_void_p]),
    "shapes_stream": (C.c_void_p, [C.c_void_p]),
    "shapes_launch_count": (C.c_int64, [C.c_void_p]),
    "shapes_version": (C.c_char_p, []),
    "shapes_last_frame_info": (C.c_int, [C.c_void_p, C.POINTER(FrameInfo)]),
    "shapes_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "shapes_stage_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "shapes_stage_name": (C.c_char_p, [C.c_int]),
}

_lib = None


def load() -> C.CDLL:
    """Load libshapes_b200.so and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m shapes_b200.build` "
            "(or __graft_entry__.build()). There is no fallback path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
