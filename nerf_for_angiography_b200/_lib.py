"""ctypes binding of libangio_b200.so (the C ABI declared in include/angio_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libangio_b200.so")

c_i32, c_i64, c_f32, c_f64, c_ptr = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_double, ctypes.c_void_p


class MlpDesc(ctypes.Structure):
    """angio_mlp_desc"""
    _fields_ = [("enc", c_i32), ("enc_basis", c_i32), ("width", c_i32), ("n_hidden", c_i32)]


class Samples(ctypes.Structure):
    """angio_samples"""
    _fields_ = [("n", c_i64), ("points", c_ptr), ("rays_o", c_ptr), ("rays_d", c_ptr), ("ray_idx", c_ptr),
                ("t_starts", c_ptr), ("t_ends", c_ptr), ("sample_idx", c_ptr), ("n_dev", c_ptr)]


_P_DESC = ctypes.POINTER(MlpDesc)
_P_SAMPLES = ctypes.POINTER(Samples)

# name -> (restype, argtypes); must list every symbol declared in include/angio_b200.h
PROTOTYPES = {
    "angio_version": (c_i32, []),
    "angio_last_error_string": (ctypes.c_char_p, []),
    "angio_sm_count": (c_i32, []),
    "angio_launch_count": (c_i64, []),
    "angio_profile_start": (c_i32, [c_ptr]),
    "angio_profile_stop": (c_i64, []),
    "angio_profile_entry": (c_i32, [c_i64, ctypes.c_char_p, c_i32, ctypes.POINTER(c_f32)]),
    "angio_sample_candidates": (c_i32, [c_ptr, c_i64, ctypes.c_uint64, c_f32, c_i32, c_ptr, c_ptr, c_ptr, c_ptr]),
    "angio_sample_rays_workspace_bytes": (c_i64, [c_i32, c_i64]),
    "angio_sample_rays": (c_i32, [c_ptr, c_i64, c_i64, ctypes.c_uint64, c_f32, c_i32, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "angio_sample_select": (c_i32, [c_ptr, c_ptr, c_i32, c_i64, ctypes.c_uint64, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "angio_raygen_flat": (c_i32, [c_ptr, c_ptr, c_i64, c_i32, c_i32, c_f64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "angio_raygen": (c_i32, [c_ptr, c_i32, c_ptr, c_ptr, c_ptr, c_i64, c_i32, c_i32, c_f64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "angio_march_runs_bytes": (c_i64, [c_i64]),
    "angio_march_count": (c_i32, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_i32, c_ptr, c_f32, c_f32, c_f32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "angio_march_head": (c_i32, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_i32, c_ptr, c_f32, c_f32, c_f32, c_i32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                  c_ptr]),
    "angio_exclusive_scan_i32": (c_i32, [c_ptr, c_i64, c_ptr, c_ptr, c_ptr]),
    "angio_march_write": (c_i32, [c_ptr, c_ptr, c_i64, c_ptr, c_i32, c_ptr, c_f32, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "angio_grid_query": (c_i32, [c_ptr, c_i64, c_ptr, c_i32, c_ptr, c_ptr, c_ptr]),
    "angio_visibility_mask": (c_i32, [c_ptr, c_ptr, c_i64, c_f32, c_f32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "angio_visibility_head_mask": (c_i32, [c_ptr, c_ptr, c_ptr, c_i64, c_i32, c_f32, c_f32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "angio_compact_head_tail": (c_i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "angio_ray_segment_counts": (c_i32, [c_ptr, c_i64, c_i32, c_i32, c_ptr, c_ptr, c_ptr]),
    "angio_ray_segment_ids": (c_i32, [c_ptr, c_ptr, c_i64, c_i32, c_ptr, c_ptr]),
    "angio_visibility_head": (c_i32, [c_ptr, c_ptr, c_i64, c_i32, c_f32, c_ptr, c_ptr]),
    "angio_compact_samples": (c_i32, [c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "angio_mlp_param_count": (c_i64, [_P_DESC]),
    "angio_mlp_input_width": (c_i32, [_P_DESC]),
    "angio_mlp_workspace_bytes": (c_i64, [_P_DESC, c_i64, c_i32, c_i32]),
    "angio_mlp_saved_bytes": (c_i64, [_P_DESC, c_i64, c_i32]),
    "angio_mlp_packed_bytes": (c_i64, [_P_DESC]),
    "angio_mlp_pack_weights": (c_i32, [_P_DESC, c_ptr, c_ptr, c_ptr]),
    "angio_mlp_forward": (c_i32, [_P_DESC, c_ptr, c_ptr, _P_SAMPLES, c_i32, c_i32, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "angio_mlp_backward": (c_i32, [_P_DESC, c_ptr, c_ptr, _P_SAMPLES, c_ptr, c_ptr, c_i32, c_ptr, c_ptr, c_i64, c_ptr]),
    "angio_composite_forward": (c_i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr]),
    "angio_composite_backward": (c_i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "angio_composite_mse_fused": (c_i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "angio_grid_cell_points": (c_i32, [c_ptr, c_ptr, c_i64, c_ptr, c_i32, c_ptr, c_ptr]),
    "angio_grid_ema_update": (c_i32, [c_ptr, c_i64, c_ptr, c_ptr, c_i64, c_f32, c_ptr, c_i64, c_ptr]),
    "angio_grid_threshold": (c_i32, [c_ptr, c_i64, c_f32, c_ptr, c_ptr, c_ptr, c_i64, c_ptr]),
    "angio_project_volume": (c_i32, [c_ptr, c_i32, c_i32, c_i32, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_i32, c_i32, c_ptr, c_ptr]),
    "angio_signal_peers": (c_i32, [c_ptr, c_i32, c_i32, ctypes.c_uint32, c_ptr]),
    "angio_adam_step_allreduce": (c_i32, [c_ptr, c_ptr, c_i32, c_ptr, ctypes.c_uint32, c_ptr, c_ptr, c_i64, c_f32, c_f32, c_f32, c_f32, c_i32,
                                          c_f32, c_i64, c_ptr, c_ptr]),
    "angio_adam_step": (c_i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_f32, c_f32, c_f32, c_f32, c_i32, c_f32, c_ptr, c_ptr]),
}

ERR_INVALID_ARG, ERR_UNSUPPORTED, ERR_WORKSPACE = -1, -2, -3
OUT_LOGIT, OUT_SIGMA, OUT_ALPHA = 0, 1, 2
PREC_FP32, PREC_BF16 = 0, 1

_lib = None


def load():
    """Load the shared library (once) and bind every prototype.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m nerf_for_angiography_b200.build` "
            "(there is no CPU / PyTorch fallback for the hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError => header/library mismatch, fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().angio_last_error_string().decode("utf-8", "replace")


def check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {last_error()}")
