"""ctypes binding of libngp_b200.so (C ABI declared in include/ngp_b200.h).

The product path has no CPU or PyTorch fallback: if the shared library is missing or a call returns a
non-zero status, a RuntimeError is raised.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_uint8, c_uint32, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NGP_B200_LIB") or os.path.join(_HERE, "lib", "libngp_b200.so")      # the override is for A/B kernel timing (tools/ab_variant.sh)

NGP_F32, NGP_F16, NGP_BF16 = 0, 1, 2
NGP_GRID_REF_ROUNDING = 1
NGP_GRID_POINT_LEVEL_KERNELS = 2

_DTYPE_ID = {torch.float32: NGP_F32, torch.float16: NGP_F16, torch.bfloat16: NGP_BF16}

_u32, _f32, _i, _p = c_uint32, c_float, c_int, c_void_p

# name -> argtypes  (restype is int status unless listed in _SPECIAL)
SIGNATURES = {
    "ngp_grid_encode_forward": [_p, _p, _p, _p, _u32, _u32, _u32, _u32, _u32, _f32, _u32, _p, _u32, _i, _u32, _i, _u32, _p],
    "ngp_grid_encode_backward": [_p, _p, _p, _p, _p, _u32, _u32, _u32, _u32, _u32, _f32, _u32, _p, _u32, _i, _u32, _i, _u32, _p],
    "ngp_grid_input_backward": [_p, _p, _p, _u32, _u32, _u32, _u32, _i, _p],
    "ngp_grid_grad_total_variation": [_p, _p, _p, _p, _f32, _u32, _u32, _u32, _u32, _f32, _u32, _u32, _i, _i, _p],
    "ngp_grid_grad_weight_decay": [_p, _p, _p, _f32, _u32, _u32, _u32, _i, _p],
    "ngp_grid_level_resolutions": [_u32, _f32, _u32, _p, _p],
    "ngp_sh_encode_forward": [_p, _p, _u32, _u32, _p, _i, _p],
    "ngp_sh_encode_backward": [_p, _p, _u32, _u32, _p, _i, _p],
    "ngp_near_far_from_aabb": [_p, _p, _p, _u32, _f32, _p, _p, _p],
    "ngp_sph_from_ray": [_p, _p, _f32, _u32, _p, _p],
    "ngp_morton3D": [_p, _u32, _p, _p],
    "ngp_morton3D_invert": [_p, _u32, _p, _p],
    "ngp_packbits": [_p, _u32, _f32, _p, _p, _p],
    "ngp_flatten_rays": [_p, _u32, _u32, _p, _p],
    "ngp_march_rays_train_count": [_p, _p, _p, _f32, _i, _f32, _u32, _u32, _u32, _u32, _p, _p, _p, _p, _p, _p, _p],
    "ngp_march_rays_train_count_aabb": [_p, _p, _p, _f32, _p, _f32, _i, _f32, _u32, _u32, _u32, _u32, _p, _u32, _p, _p, _p, _p, _p, _p],
    "ngp_march_rays_train_write": [_p, _p, _p, _p, _f32, _i, _f32, _u32, _u32, _u32, _u32, _p, _p, _p, _p, _u32, _p, _p, _p, _p, _p, _p, _p],
    "ngp_composite_train_mse": [_p, _p, _p, _p, _u32, _p, _u32, _f32, _f32, _p, _f32, _p, _p, _p, _p, _p, _p, _i, _p, _p],
    "ngp_composite_train_loss": [_p, _p, _p, _p, _u32, _p, _u32, _f32, _f32, _p, _f32, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p],
    "ngp_march_rays_train_count_ex": [_p, _p, _p, _f32, _p, _p, _p, _f32, _i, _f32, _u32, _u32, _u32, _u32, _p, _u32, _p, _p, _p, _p, _p, _p],
    "ngp_adaptive_num_rays": [_p, _p, _u32, _u32, _p],
    "ngp_uniform": [_p, _u32, c_uint64, _p, _p],
    "ngp_composite_rays_train_forward": [_p, _p, _p, _p, _u32, _u32, _f32, _p, _p, _p, _p, _p],
    "ngp_composite_rays_train_backward": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _u32, _u32, _f32, _p, _p, _p],
    "ngp_march_rays_train_backward": [_p, _p, _p, _p, _u32, _u32, _p, _p, _p],
    "ngp_march_rays": [_u32, _u32, _p, _p, _p, _p, _f32, _i, _f32, _u32, _u32, _u32, _p, _p, _p, _p, _p, _p, _p, _p],
    "ngp_composite_rays": [_u32, _u32, _f32, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "ngp_compact_rays_alive": [_p, _u32, _p, _p, _p, _p],
    "ngp_march_rays_dev": [_p, _u32, _p, _p, _p, _p, _f32, _i, _f32, _u32, _u32, _u32, _p, _p, _p, _p, _p, _p, _p],
    "ngp_composite_rays_dev": [_p, _u32, _f32, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "ngp_compact_rays_alive_dev": [_p, _p, _u32, _u32, _u32, _u32, _p, _p, _p, _p, _p],
    "ngp_occ_sample_positions": [_p, _p, _u32, _u32, _f32, _p, _p, _p],
    "ngp_occ_scatter_sigmas": [_p, _p, _u32, _p, _p],
    "ngp_occ_ema_update": [_p, _p, _u32, _f32, _p, _p, _p],
    "ngp_occ_sample_partial": [_p, _u32, _f32, _p, _u32, _p, _p, _p, _p, _p, _p],
    "ngp_mark_untrained_grid": [_p, _p, _u32, _u32, _p, _u32, _p, _f32, _p, _u32, _u32, _f32, _p],
    "ngp_mlp_forward": [_p, _u32, _p, _p, _u32, _u32, _i, _p, _u32, _p, _p],
    "ngp_mlp_backward": [_p, _u32, _p, _u32, _p, _p, _p, _u32, _u32, _i, _p, _u32, _p, _p],
    "ngp_field_forward_density": [_p, _p, _p, _p, _p, _p, _f32, _f32, _u32, _u32, _u32, _i, _u32, _p, _p, _u32, _u32, _p, _i, _f32, _p, _p, _p, _p, _u32, _p],
    "ngp_field_backward_density": [_p, _p, _p, _p, _u32, _p, _p, _p, _p, _f32, _f32, _u32, _u32, _u32, _i, _u32, _p, _p, _p, _u32, _u32, _p, _i, _f32, _p, _p, _p, _p],
    "ngp_sh_dirs_backward": [_p, _u32, _u32, _p, _u32, _p, _p, _p],
    "ngp_field_backward_full": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _f32, _f32, _u32, _u32, _u32, _i, _u32, _p, _p, _p, _p, _u32, _p, _i, _f32, _i, _p, _p, _p, _p, _p, _p, _p, _p],
    "ngp_field_forward_full": [_p, _p, _p, _p, _p, _p, _f32, _f32, _u32, _u32, _u32, _i, _u32, _p, _p, _p, _p, _u32, _p, _i, _f32, _i, _p, _p, _p, _p, _p, _p, _p, _p],
    "ngp_mlp_forward_rgb": [_p, _u32, _p, _p, _u32, _u32, _p, _i, _i, _p, _p, _p],
    "ngp_mlp_backward_rgb": [_p, _p, _i, _p, _u32, _p, _p, _p, _u32, _u32, _p, _i, _p, _u32, _p, _p],
    "ngp_fused_adam": [_p, _p, _i, _p, _i, _p, _p, c_uint64, _f32, _f32, _f32, _f32, _f32, _u32, _p, _p, _p, _p, _i, _p],
    "ngp_diag_l2_rate": [_p, _u32, _u32, _u32, _i, _p, _p],
    "ngp_freq_encode_forward": [_p, _u32, _u32, _u32, _u32, _p, _p],
    "ngp_freq_encode_backward": [_p, _p, _u32, _u32, _u32, _u32, _p, _p],
    "ngp_pose_rays_forward": [_p, _p, _u32, _p, _p, _u32, _u32, _p, _p, _p],
    "ngp_pose_rays_backward": [_p, _p, _p, _p, _u32, _p, _p, _u32, _u32, _p, _p],
    "ngp_adam_step_counter": [_p, _p, _p],
    "ngp_grad_scaler_update": [_p, _p, _p, _p, _p, _f32, _f32, _i, _u32, _p],
    "ngp_dp_fused_adam": [_p, _i, _p, _i, _u32, _u32, _p, _p, _p, c_uint64, c_uint64, _f32, _f32, _f32, _f32, _f32, _p, _p, _p, _p, _p, _u32, _p],
    "ngp_dp_check_publish": [_p, _p, _p, _u32, _p, _p, _p, _u32, _u32, _p],
    "ngp_dp_finish": [_p, _u32, _p, _p, _p, c_uint64, _p, c_uint64, _p],
    "ngp_dp_publish_flag": [_p, _p, _u32, _u32, _p],
    "ngp_dp_merge_flags": [_p, _u32, _p, _p],
    "ngp_check_finite": [_p, _i, c_uint64, _p, _p],
    "ngp_small_adam": [_p, _p, _p, _p, _u32, _f32, _f32, _f32, _f32, _f32, _p, _p, _p, _p, _p],
    "ngp_dp_small_adam": [_p, _u32, _p, _p, _p, _u32, _f32, _f32, _f32, _f32, _f32, _p, _p, _p, _p, _p],
    "ngp_check_finite_multi": [_p, _p, _p, _u32, _p, _p, _p, _p],
}
_SPECIAL = {
    "ngp_abi_version": ([], c_int),
    "ngp_status_string": ([c_int], c_char_p),
    "ngp_last_cuda_error": ([], c_char_p),
}



class LossOpts(ctypes.Structure):
    """ngp_loss_opts of include/ngp_b200.h (optional terms of ngp_composite_train_loss)."""
    _fields_ = [("bg_rays", _p), ("target_alpha", _p), ("lossmult", _p), ("loss_weight", _p), ("inv_norm_dev", _p), ("n_rays_dev", _p),
                ("lambda_entropy", _f32), ("entropy_ray", _p), ("weights_sum_out", _p), ("depth_out", _p), ("parts_out", _p), ("loss_scale_dev", _p)]


_lib = None


def exported_symbols():
    """Every symbol include/ngp_b200.h declares."""
    return sorted(list(SIGNATURES) + list(_SPECIAL))


def load():
    """Loads the shared library (once).  Raises RuntimeError if it is absent -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"raw_ngp_b200: {LIB_PATH} not found. Build it with `python -m raw_ngp_b200.build` "
            "(nvcc, sm_100a). There is no CPU / PyTorch fallback for these operators.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
    for name, (argtypes, restype) in _SPECIAL.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def dtype_id(dtype):
    try:
        return _DTYPE_ID[dtype]
    except KeyError:
        raise RuntimeError(f"raw_ngp_b200: unsupported dtype {dtype} (float32, float16, bfloat16 only)")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("raw_ngp_b200: operator called with a CPU tensor; these operators are CUDA-only (sm_100a)")


weights_epoch = 0  # bumped whenever a native optimizer may have written parameters behind torch's version counters
launch_count = 0  # kernels launched through the C ABI by this process (every entry point is one launch)


def call(name, *args):
    """Calls an entry point on the current stream's device and raises on a non-zero status."""
    global launch_count
    lib = load()
    launch_count += 1
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.ngp_status_string(rc).decode()
        if rc == -5:
            msg += ": " + lib.ngp_last_cuda_error().decode()
        raise RuntimeError(f"raw_ngp_b200.{name} failed ({rc}): {msg}")
