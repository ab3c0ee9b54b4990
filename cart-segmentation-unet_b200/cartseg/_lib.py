"""ctypes binding of libcartseg.so — the C ABI declared in include/cartseg.h.

There is no fallback of any kind: if the shared library is missing or a call fails, an exception
is raised (the reference's only native op does the same for CPU tensors,
src/training/abl_training/losses/lsr_cpp/csrc/lsr_kernel.cu:300-302).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CARTSEG_LIB_PATH") or os.path.join(_HERE, "libcartseg.so")   # override: A/B runs of two builds

NUM_PARAMS = 82
NUM_BN = 18
NUM_BWD_STAGES = 23


class UnetTensors(C.Structure):
    _fields_ = [
        ("param", C.c_void_p * NUM_PARAMS),
        ("grad", C.c_void_p * NUM_PARAMS),
        ("running_mean", C.c_void_p * NUM_BN),
        ("running_var", C.c_void_p * NUM_BN),
        ("num_batches_tracked", C.c_void_p * NUM_BN),
    ]


class LossDesc(C.Structure):
    _fields_ = [
        ("rows", C.c_int),
        ("n", C.c_longlong),
        ("w_elem", C.c_float), ("alpha", C.c_float), ("gamma", C.c_float),
        ("elem_sum", C.c_int),
        ("w_dice", C.c_float), ("smooth", C.c_float),
        ("w_bgt", C.c_float), ("w_bpred", C.c_float),
        ("use_abs", C.c_int),
        ("per_row", C.c_int),
    ]


ABL_LADDER = 80


class AblDesc(C.Structure):
    _fields_ = [
        ("batch", C.c_int), ("height", C.c_int), ("width", C.c_int),
        ("max_n", C.c_float), ("label_smoothing", C.c_float), ("max_clip_dist", C.c_float),
        ("ignore_label", C.c_longlong),
        ("per_image_maps", C.c_int),
        ("eps_ladder", C.c_float * ABL_LADDER),
    ]


class ImageDesc(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("height", C.c_int), ("width", C.c_int), ("pitch", C.c_int),
        ("canvas_h", C.c_int), ("canvas_w", C.c_int),
        ("x0", C.c_int), ("y0", C.c_int),
        ("reserved", C.c_int),
    ]


_P = C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "cs_last_error": (C.c_char_p, []),
    "cs_version": (C.c_int, []),
    "cs_kernel_launch_count": (C.c_longlong, []),
    "cs_unet_plan_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "cs_unet_plan_destroy": (None, [_P]),
    "cs_unet_plan_workspace_bytes": (C.c_size_t, [_P]),
    "cs_unet_plan_bind": (C.c_int, [_P, _P, C.c_size_t]),
    "cs_unet_pack_weights": (C.c_int, [_P, C.POINTER(UnetTensors), _P]),
    "cs_unet_forward": (C.c_int, [_P, C.POINTER(UnetTensors), _P, C.c_int, _P, _P]),
    "cs_unet_backward": (C.c_int, [_P, C.POINTER(UnetTensors), _P, C.c_int, C.c_int, C.c_int, _P]),
    "cs_unet_set_overlap": (C.c_int, [_P, C.c_int]),
    "cs_unet_set_deferred_join": (C.c_int, [_P, C.c_int]),
    "cs_unet_plan_set_sm_limit": (C.c_int, [_P, C.c_int]),
    "cs_unet_backward_wait": (C.c_int, [_P, _P]),
    "cs_unet_backward_held_stages": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]),
    "cs_unet_profile": (C.c_int, [_P, C.c_int]),
    "cs_unet_profile_read": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "cs_unet_trace": (C.c_int, [_P, C.c_int]),
    "cs_unet_trace_read": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "cs_unet_debug_read": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_int), _P, _P]),
    "cs_unet_stage_params": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.c_int]),
    "cs_sdf_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "cs_sdf": (C.c_int, [_P, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P, _P]),
    "cs_loss_scratch_bytes": (C.c_size_t, [C.c_int]),
    "cs_loss_forward": (C.c_int, [C.POINTER(LossDesc), _P, _P, _P, _P, _P, _P, _P]),
    "cs_loss_backward": (C.c_int, [C.POINTER(LossDesc), _P, _P, _P, _P, _P, _P, _P, _P]),
    "cs_focal_map_forward": (C.c_int, [_P, _P, C.c_longlong, C.c_float, C.c_float, _P, _P]),
    "cs_focal_map_backward": (C.c_int, [_P, _P, _P, C.c_longlong, C.c_float, C.c_float, _P, _P]),
    "cs_threshold_stats": (C.c_int, [_P, _P, C.c_int, C.c_longlong, _P, C.c_int, _P, _P, _P]),
    "cs_threshold_mask": (C.c_int, [_P, C.c_longlong, C.c_float, _P, _P]),
    "cs_abl_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "cs_abl_forward": (C.c_int, [C.POINTER(AblDesc), _P, _P, _P, _P, _P]),
    "cs_abl_backward": (C.c_int, [C.POINTER(AblDesc), _P, _P, _P, _P, _P]),
    "cs_abl_debug_read": (C.c_int, [C.POINTER(AblDesc), _P, C.POINTER(C.c_float), C.POINTER(C.c_int),
                                    C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), _P, _P, _P]),
    "cs_ensemble_accumulate": (C.c_int, [_P, C.c_float, C.c_longlong, C.c_int, _P, _P]),
    "cs_pseudo_qc": (C.c_int, [_P, C.c_int, C.c_longlong, C.c_float, C.c_int, _P, _P, _P]),
    "cs_mask_cleanup_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "cs_mask_cleanup": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "cs_letterbox_geometry": (C.c_int, [C.c_int, C.c_int, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                        C.POINTER(C.c_int)]),
    "cs_preproc_images": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, _P, _P]),
    "cs_preproc_masks": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "cs_layer_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "cs_conv3x3_fprop": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P, _P, _P, _P, _P]),
    "cs_conv3x3_fprop_bnrelu": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P, _P, _P, _P, _P]),
    "cs_conv3x3_wgrad_bnrelu": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "cs_conv3x3_dgrad": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P, _P, _P]),
    "cs_conv3x3_wgrad": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "cs_convT2x2_fprop": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, _P, C.c_int, _P, _P]),
    "cs_convT2x2_dgrad": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P, _P, _P]),
    "cs_convT2x2_wgrad": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class CartsegError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """The loaded library.  Raises if it has not been built (python __graft_entry__.py / make -C csrc)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CartsegError(
                f"{LIB_PATH} is missing: build it with `make -C cart-segmentation-unet_b200/csrc` "
                "(or __graft_entry__.build()).  cartseg has no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            if os.environ.get("CARTSEG_LIB_PATH") and not hasattr(handle, name):
                continue                        # an older build loaded for an A/B run may lack newer developer hooks
            fn = getattr(handle, name)          # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().cs_last_error()
        raise CartsegError(f"{what} failed ({code}): {msg.decode() if msg else 'unknown error'}")


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int | None:
    """Device pointer of a CUDA tensor (None passes NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise CartsegError("cartseg ops take CUDA tensors only (no CPU fallback)")
    return t.data_ptr()
