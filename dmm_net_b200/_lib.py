"""ctypes binding of libdmm_b200.so (C ABI declared in include/dmm_b200.h).

There is NO CPU or pure-torch fallback: if the library is missing or a tensor is not on a CUDA device the
ops raise.  ``load()`` only dlopens the library (works without a GPU, used by the CPU-side symbol test).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_size_t, c_void_p, POINTER

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libdmm_b200.so")

_vp, _i, _f, _ll, _sz = c_void_p, c_int, c_float, c_longlong, c_size_t

# name -> (restype, argtypes); mirrors include/dmm_b200.h one to one
SIGNATURES = {
    "dmm_b200_version": (_i, []),
    "dmm_b200_arch": (c_char_p, []),
    "dmm_b200_error_string": (c_char_p, [_i]),
    "dmm_b200_last_cuda_error": (_i, []),
    "dmm_b200_limits": (_i, [POINTER(c_int)]),
    "dmm_mask_iou_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "dmm_mask_iou_pairwise": (_i, [_vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp,
                                   _vp, _vp, _sz, _vp]),
    "dmm_mask_iou_pairwise_ptrs": (_i, [_vp, _i, _vp, _ll, _vp, _ll, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _f, _vp,
                                        _vp, _vp, _sz, _vp]),
    "dmm_packed_words": (_ll, [_ll]),
    "dmm_host_pack_masks": (_i, [_vp, _ll, _ll, _vp, _i]),
    "dmm_host_pack_masks2": (_i, [_vp, _ll, _vp, _vp, _ll, _vp, _ll, _i]),
    "dmm_host_read_bandwidth": (_i, [_vp, _ll, _i, _i, POINTER(ctypes.c_double)]),
    "dmm_mask_pack_bits": (_i, [_vp, _ll, _i, _vp, _vp]),
    "dmm_mask_iou_packed_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "dmm_mask_iou_pairwise_packed": (_i, [_vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _f,
                                          _vp, _vp, _vp, _sz, _vp]),
    "dmm_mask_iou_rowwise_workspace_bytes": (_sz, [_i, _i]),
    "dmm_mask_iou_rowwise": (_i, [_vp, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    "dmm_cosine_pairwise": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _f, _vp, _vp]),
    "dmm_cosine_pairwise_impl": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _f, _vp, _i, _vp]),
    "dmm_cosine_pairwise_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _f, _vp, _vp, _vp]),
    "dmm_relax_saved_bytes": (_sz, [_i, _i, _i]),
    "dmm_relax_solve": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _i, _f, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                             _vp, _vp, _vp, _vp]),
    "dmm_relax_solve_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _i, _i,
                                 _f, _i, _i, _vp, _vp, _vp]),
    "dmm_assign_apply": (_i, [_vp, _vp, _ll, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _ll, _vp]),
    "dmm_assign_apply_ptrs": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _ll, _vp]),
    "dmm_assign_apply_bwd_ptrs": (_i, [_vp, _ll, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dmm_assign_apply_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "dmm_assign_apply_bwd": (_i, [_vp, _ll, _vp, _ll, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz,
                                  _vp]),
    "dmm_roi_mean_pool_workspace_bytes": (_sz, [POINTER(c_int), POINTER(c_int), _i, _i, _i]),
    "dmm_roi_mean_pool": (_i, [POINTER(c_void_p), POINTER(c_int), POINTER(c_int), _i, _i, _vp, _i, _vp, _vp, _sz, _i, _vp]),
    "dmm_roi_mean_pool_bwd_workspace_bytes": (_sz, [POINTER(c_int), POINTER(c_int), _i, _i, _i]),
    "dmm_roi_mean_pool_bwd": (_i, [_vp, POINTER(c_int), POINTER(c_int), _i, _i, _vp, _i, POINTER(c_void_p), _vp, _sz, _i,
                                   POINTER(c_int), _vp]),
    "dmm_mask_pyramid_level_size": (_i, [_i, _i, _i, POINTER(c_int), POINTER(c_int)]),
    "dmm_mask_pyramid": (_i, [_vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _i, _i, POINTER(c_void_p), _vp]),
    "dmm_mask_pyramid_bwd": (_i, [POINTER(c_void_p), _vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "dmm_merge_labels": (_i, [_vp, _ll, _i, _i, _i, _vp, _vp, _vp]),
    "dmm_paste_masks_workspace_bytes": (_sz, [_i]),
    "dmm_paste_masks": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dmm_paste_apply": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _ll, _vp]),
    "dmm_paste_apply_bwd": (_i, [_vp, _ll, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp]),
    "dmm_box_nms": (_i, [_vp, _vp, _vp, _i, _i, _f, _i, _vp, _vp, _vp]),
}

_lib = None


def load() -> ctypes.CDLL:
    """dlopen the library and attach the signatures.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: dmm_net_b200 has no CPU / pure-torch fallback. "
            "Build it with `python -m dmm_net_b200.build` (needs nvcc, targets sm_100a).")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    if lib.dmm_b200_version() != 0x000200:
        raise RuntimeError("libdmm_b200.so version mismatch: rebuild with `python -m dmm_net_b200.build --force`")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        lib = load()
        msg = lib.dmm_b200_error_string(rc).decode()
        extra = f" (cudaError {lib.dmm_b200_last_cuda_error()})" if rc == 4 else ""
        raise RuntimeError(f"{what} failed: {msg}{extra}")
