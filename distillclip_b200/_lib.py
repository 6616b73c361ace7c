"""ctypes binding of `libdistillclip_b200.so` (C ABI in include/distillclip_b200.h).

There is deliberately NO fallback: if the library is missing or a call fails, we raise.  The product
path never routes through torch eager ops or the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libdistillclip_b200.so")

BF16, F16, F32 = 0, 1, 2
MAX_LAYERS, MAX_PARTIALS, MAX_TERMS = 16, 4096, 16
TOWER_MAX_SEG, TOWER_MAX_TERMS = 40, 8

_vp, _i64p, _i32p, _f32p = C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_float)
_vpp = C.POINTER(C.c_void_p)

# name -> argtypes ; every function returns int (0 = ok) unless listed in _SPECIAL
SIGNATURES = {
    "dcb_mse_fwd_bwd": [C.c_int, _vpp, _vpp, _vpp, _i64p, C.c_int, C.c_int, C.c_int, C.c_float, _vp,
                        C.POINTER(C.c_int), _vp],
    "dcb_attn_kl_fwd_bwd": [C.c_int, _vpp, _vpp, _vpp, _i64p, _i32p, _i32p, _i64p, C.c_int, C.c_int, C.c_int,
                            C.c_float, _vp, C.POINTER(C.c_int), _vp],
    "dcb_finalize": [C.c_int, _vpp, _i32p, _f32p, _f32p, _vp, _vp],
    "dcb_tower_fwd_bwd": [C.c_int, _i32p, _i32p, _vpp, _vpp, _vpp, _i64p, _i64p, _i32p, _i32p, _i64p, _i32p, _f32p,
                          C.c_int, _f32p, _f32p, C.c_int, C.c_int, _vp, C.c_int, C.c_uint32, C.c_int, _vp, _vp, _vp, _vpp,
                          _f32p, _f32p, _vp],
    "dcb_row_inv_norm": [C.c_int, _vpp, _vpp, _i64p, C.c_int64, C.c_int, _vp],
    "dcb_transpose_norm_f16": [_vp, _vp, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_int, _vp],
    "dcb_clip_row_stats": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                           C.c_int, C.c_float, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "dcb_clip_rank_counts": [_vp, _vp, _vp, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, _vp, _vp, _vp, _vp, _vp],
    "dcb_clip_col_finish": [_vp, C.c_int64, _vp, C.c_int64, C.c_int64, C.c_float, C.c_int, _vp, _vp, _vp],
    "dcb_clip_losses": [_vp, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_int, _vp, _vp, _vp],
    "dcb_clip_grad_coef": [_vp, C.c_int64, C.c_int64, C.c_float, C.c_int, _vp, _vp, _vp, _vp],
    "dcb_clip_row_grads": [_vp, _vp, _vp, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64,
                           C.c_int64, C.c_int64, C.c_int, C.c_float, _vp, _vp],
    "dcb_clip_row_grads_pair": [_vp, _vp, _vp, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64,
                                C.c_int64, C.c_int64, C.c_int, C.c_float, _vp, _vp, C.c_int64, _vp, _vp, _vp],
    "dcb_clip_col_grads_from_g": [_vp, C.c_int64, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _vp, _vp],
    "dcb_clip_col_grads_scatter": [_vp, C.c_int64, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _vp, C.c_int, C.c_int, _vp],
    "dcb_clip_grad_finish": [_vp, C.c_int, _vp, _vp, _vp, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                             _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp],
    "dcb_memcpy_async": [_vp, _vp, C.c_int64, _vp],
    "dcb_peer_gather": [C.c_int, _vpp, _vpp, _i64p, _vp],
    "dcb_clip_prep": [C.c_int, _vpp, _vpp, _vpp, _vpp, _i64p, C.c_int64, C.c_int64, C.c_int, _vp],
    "dcb_clip_fwd_chunk": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                           C.c_float, _vp, _vp, _vp, C.c_int64, _vp, _vp, _vp],
    "dcb_clip_post1": [_vp, C.c_int, _vp, _vp, C.c_int, C.c_int64, C.c_int64, C.c_float, C.c_int, C.c_int64, _vp, _vp, _vpp,
                       C.c_int, _vp, _vp, _vp, _vp],
    "dcb_clip_post2": [_vp, C.c_int, C.c_int64, C.c_int64, C.c_float, C.c_int, C.c_int64, _f32p, _vp, _vp, _vp, _vp, _vp, _vp],
    "dcb_clip_pair_bwd": [_vp, _vp, _vp, _vp, _vp, C.c_int64, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vpp, _f32p,
                          C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_float,
                          _vp, _vp, C.c_int64, _vp],
    "dcb_clip_g_tiles": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vpp, _f32p, C.c_int, C.c_int64, C.c_int64,
                         C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_float, _vp, C.c_int64, _vp],
    "dcb_clip_row_grads_from_g": [_vp, C.c_int64, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _vp, _vp],
    "dcb_clip_finish2": [_vp, C.c_int, C.c_int64, _vp, _vp, _vp, C.c_int64, _vp, _vp, C.c_int64, C.c_int64,
                         _vp, C.c_int, C.c_int64, _vp, _vp, _vp, C.c_int64, _vp, _vp, C.c_int64, C.c_int64,
                         C.c_int64, C.c_int64, _vpp, _f32p, _vp, _vp, C.c_int, C.c_int, _vp],
    "dcb_value_map_kl_fwd_bwd": [_vp, _vp, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_float, _vp,
                                 C.POINTER(C.c_int), _vp, _vp, C.c_float, _vp],
    "dcb_row_softmax_stats": [_vp, _vp, C.c_int64, C.c_int64, C.c_int, C.c_float, C.c_int, _vp, _vp, _vp],
    "dcb_row_softmax_grads": [_vp, _vp, C.c_int64, C.c_int64, C.c_int, C.c_float, C.c_int, _vp, _vp, _vp, C.c_int, _vp],
    "dcb_logits_row_stats": [_vp, C.c_int64, C.c_int64, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_float,
                             C.c_int, _vp, _vp, _vp],
    "dcb_logits_row_grads": [_vp, C.c_int64, C.c_int64, _vp, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_float,
                             C.c_int, _vp, _vp, _vp, C.c_int, _vp],
}
_SPECIAL = {
    "dcb_clip_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int64]),
    "dcb_clip_grad_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int64, C.c_int64]),
    "dcb_clip_grad_splits": (C.c_int, [C.c_int64, C.c_int64, C.c_int64]),
    "dcb_clip_pair_splits": (C.c_int, [C.c_int64, C.c_int64, C.c_int64]),
    "dcb_clip_pair_supported": (C.c_int, [C.c_int64]),
    "dcb_clip_gt_splits": (C.c_int, [C.c_int64, C.c_int64, C.c_int64]),
    "dcb_clip_rg_splits": (C.c_int, [C.c_int64, C.c_int64, C.c_int64]),
    "dcb_clip_gt_splits_scatter": (C.c_int, [C.c_int64, C.c_int64, C.c_int64]),
    "dcb_clip_fwd_chunk_parts": (C.c_int, [C.c_int64, C.c_int64]),
    "dcb_clip_slot_floats": (C.c_int64, [C.c_int64, C.c_int64]),
    "dcb_clip_post_scratch_bytes": (C.c_int64, [C.c_int64, C.c_int64]),
    "dcb_tower_grid": (C.c_int, []),
    "dcb_version": (C.c_int, []),
    "dcb_compiled_arch": (C.c_int, []),
    "dcb_last_error": (C.c_char_p, []),
}
EXPORTED = sorted(list(SIGNATURES) + list(_SPECIAL))

_lib = None
_lock = threading.Lock()


class DistillClipB200Error(RuntimeError):
    pass


def load():
    """Load the CUDA library (once). Raises DistillClipB200Error if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise DistillClipB200Error(
                f"{LIB_PATH} is missing: build it with `python -m distillclip_b200.build` "
                "(or __graft_entry__.build()). There is no CPU or eager fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes, fn.restype = args, C.c_int
        for name, (res, args) in _SPECIAL.items():
            fn = getattr(lib, name)
            fn.argtypes, fn.restype = args, res
        _lib = lib
    return _lib


# kernels launched per C-ABI call (bench.py reports the sum as "gpu_launches")
_LAUNCHES_PER_CALL = {"dcb_clip_row_stats": 3, "dcb_clip_rank_counts": 2, "dcb_memcpy_async": 0}
LAUNCHES = 0


def call(name: str, *args) -> None:
    global LAUNCHES
    lib = load()
    rc = getattr(lib, name)(*args)
    LAUNCHES += _LAUNCHES_PER_CALL.get(name, 1)
    if rc != 0:
        raise DistillClipB200Error(f"{name} failed: {lib.dcb_last_error().decode(errors='replace')}")


def ptr_array(ptrs):
    return (C.c_void_p * len(ptrs))(*[C.c_void_p(int(p)) if p else None for p in ptrs])


def i64_array(v):
    return (C.c_int64 * len(v))(*[int(x) for x in v])


def i32_array(v):
    return (C.c_int32 * len(v))(*[int(x) for x in v])


def f32_array(v):
    return (C.c_float * len(v))(*[float(x) for x in v])
