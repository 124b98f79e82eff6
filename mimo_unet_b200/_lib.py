"""ctypes binding of libmimo_b200.so (the C ABI declared in include/mimo_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  If the shared object is missing
the import of any compute entry point raises, loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmimo_b200.so")
CSRC = os.path.join(_HERE, "csrc")


class MimoError(RuntimeError):
    pass


class Act(C.Structure):
    """mimo_act_t"""
    _fields_ = [("ptr", C.c_void_p), ("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("pad", C.c_int),
                ("cpitch", C.c_int), ("c_off", C.c_int), ("c", C.c_int)]


class UnetConfig(C.Structure):
    """mimo_unet_config_t"""
    _fields_ = [("in_channels", C.c_int), ("out_channels", C.c_int), ("num_subnetworks", C.c_int),
                ("filter_base_count", C.c_int), ("batch", C.c_int), ("height", C.c_int), ("width", C.c_int)]


def build(verbose: bool = False) -> str:
    """Compiles the library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise MimoError("building libmimo_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


_lib: Optional[C.CDLL] = None

vp, i32, i64, f32, f64, sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double, C.c_size_t
ActP = C.POINTER(Act)

# name -> (restype, argtypes); must list every symbol include/mimo_b200.h declares
SIGNATURES = {
    "mimo_version": (i32, []),
    "mimo_last_error": (C.c_char_p, []),
    "mimo_check_device": (i32, []),
    "mimo_pack_input": (i32, [vp, i64, i64, vp, Act, vp]),
    "mimo_weight_pack": (i32, [vp, i32, i32, vp, i32, vp, i32, vp]),
    "mimo_conv3x3_m_tiles": (i32, [i32, i32, i32]),
    "mimo_conv3x3": (i32, [Act, i32, vp, i32, i32, vp, i32, vp, vp, vp, i32, vp]),
    "mimo_conv3x3_wgrad": (i32, [Act, Act, vp, i32, vp, i32, vp]),
    "mimo_bn_finalize": (i32, [vp, vp, i32, i32, i32, f64, vp, vp, vp, vp, vp, vp, f32, f32, vp, vp, vp, vp, vp]),
    "mimo_bn_eval_affine": (i32, [i32, vp, vp, vp, vp, vp, f32, vp, vp, vp, vp, vp]),
    "mimo_bn_relu_apply": (i32, [vp, i32, vp, vp, vp, Act, ActP, vp]),
    "mimo_maxpool2x2": (i32, [Act, Act, vp, vp]),
    "mimo_upsample_bilinear2x": (i32, [Act, Act, vp]),
    "mimo_upsample_bilinear2x_bwd": (i32, [Act, Act, i32, vp]),
    "mimo_upsample_concat": (i32, [Act, Act, Act, vp]),
    "mimo_maxunpool2x2": (i32, [Act, vp, Act, vp]),
    "mimo_convtranspose2x2": (i32, [Act, vp, vp, Act, vp]),
    "mimo_unpack_nchw": (i32, [Act, vp, vp]),
    "mimo_grad_gather": (i32, [ActP, ActP, ActP, Act, i32, vp]),
    "mimo_bn_bwd_scratch_floats": (sz, [i32]),
    "mimo_bn_relu_bwd": (i32, [Act, vp, i32, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, i32, Act, vp]),
    "mimo_bn_relu_bwd_folded": (i32, [Act, Act, vp, i32, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, i32, Act, vp]),
    "mimo_wgrad_streamk_schedule": (i32, [i32, i32, i64, i32, i32, vp, i32]),
    "mimo_mask_mul": (i32, [Act, vp, i32, f32, vp]),
    "mimo_unet_set_elementwise_dropout": (i32, [vp, vp, f32, vp, f32]),
    "mimo_head1x1": (i32, [Act, vp, vp, i32, vp, i64, vp]),
    "mimo_head1x1_bwd_scratch_floats": (sz, [i32, i32]),
    "mimo_head1x1_bwd": (i32, [Act, vp, i32, vp, i64, vp, Act, vp, vp, vp, i32, vp]),
    "mimo_laplace_scratch_floats": (sz, []),
    "mimo_laplace_nll_fwd": (i32, [vp, i64, vp, i64, vp, i64, vp, i64, i64, i64, f32, f32, vp, vp, vp, vp]),
    "mimo_laplace_nll_bwd": (i32, [vp, i64, vp, i64, vp, i64, vp, i64, i64, i64, f32, f32, vp, i32, f32, vp, vp, vp]),
    "mimo_lossbuffer_bytes": (sz, [i32, i32]),
    "mimo_lossbuffer_init": (i32, [vp, i32, i32, f32, vp]),
    "mimo_lossbuffer_get_weights": (i32, [vp, vp, vp]),
    "mimo_lossbuffer_add": (i32, [vp, vp, vp]),
    "mimo_laplace_train_scratch_floats": (sz, [i32, i32, i32, i64]),
    "mimo_laplace_nll_train": (i32, [vp, vp, i64, i64, vp, i64, i64, vp, i32, i32, i32, i64, f32, f32, vp, vp, i32,
                                     vp, vp, vp, vp, vp, vp]),
    "mimo_laplace_train_metrics_scratch_floats": (sz, [i32, i32, i32, i64]),
    "mimo_laplace_nll_train_metrics": (i32, [vp, vp, i64, i64, vp, i64, i64, vp, i32, i32, i32, i64, f32, f32, vp, vp, i32,
                                             vp, vp, vp, vp, vp, vp, vp]),
    "mimo_gaussian_nll_fwd": (i32, [vp, i64, vp, i64, vp, i64, vp, i64, i64, i64, f32, f32, vp, vp, vp, vp]),
    "mimo_gaussian_nll_bwd": (i32, [vp, i64, vp, i64, vp, i64, vp, i64, i64, i64, f32, f32, vp, i32, f32, vp, vp, vp]),
    "mimo_gaussian_nll_train_metrics": (i32, [vp, vp, i64, i64, vp, i64, i64, vp, i32, i32, i32, i64, f32, f32, vp, vp, i32,
                                              vp, vp, vp, vp, vp, vp, vp]),
    "mimo_evidential_head": (i32, [vp, vp, i64, i64, vp]),
    "mimo_evidential_head_bwd": (i32, [vp, vp, vp, i64, i64, vp]),
    "mimo_evidential_loss_fwd": (i32, [vp, vp, vp, i64, i64, vp, vp, vp, vp]),
    "mimo_evidential_loss_bwd": (i32, [vp, vp, vp, i64, i64, vp, i32, f32, vp, vp]),
    "mimo_scale_by_scalar": (i32, [vp, i64, vp, vp]),
    "mimo_ensemble_aggregate": (i32, [vp, i64, i64, vp, i64, i64, i32, i32, i64, vp, vp, vp, vp]),
    "mimo_validation_scratch_floats": (i32, [i32]),
    "mimo_validation_laplace": (i32, [vp, vp, i64, i64, vp, vp, i32, i32, i64, f32, f32, vp, vp, vp, vp, vp, vp, vp]),
    "mimo_unet_plan_create": (i32, [C.POINTER(UnetConfig), C.POINTER(vp)]),
    "mimo_unet_plan_destroy": (None, [vp]),
    "mimo_unet_workspace_bytes": (sz, [vp]),
    "mimo_unet_num_state": (i32, [vp]),
    "mimo_unet_num_double_convs": (i32, [vp]),
    "mimo_unet_dropout_channels": (i32, [vp, i32]),
    "mimo_unet_bind": (i32, [vp, vp, sz, C.POINTER(vp), C.POINTER(vp), i32]),
    "mimo_unet_forward": (i32, [vp, vp, vp, i32, C.POINTER(vp), vp, vp]),
    "mimo_unet_backward": (i32, [vp, vp, vp, vp, i32, vp]),
    "mimo_unet_debug_view": (i32, [vp, C.c_char_p, ActP, C.POINTER(i32)]),
    "mimo_unet_stack_layout": (i32, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mimo_unet_set_inference_fusion": (i32, [vp, i32]),
    "mimo_unet_set_backward_events": (i32, [vp, C.POINTER(C.c_void_p)]),
    "mimo_unet_backward_stage_first_state": (i32, [vp, i32]),
    "mimo_unet_last_launches": (i32, [vp]),
    "mimo_unet_graph_state": (i32, [vp]),
    "mimo_unet_profile_classes": (i32, []),
    "mimo_unet_profile_class_name": (C.c_char_p, [i32]),
    "mimo_unet_profile_enable": (i32, [vp, i32]),
    "mimo_unet_profile_read": (i32, [vp, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "mimo_unet_profile_read_launches": (i32, [vp, i32, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mimo_unet_profile_read_launches_ex": (i32, [vp, i32, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mimo_conv_kernel_name": (C.c_char_p, [i32]),
    "mimo_unet_node_name": (C.c_char_p, [vp, i32]),
}


def lib() -> C.CDLL:
    """Loads (once) and returns the shared library; raises MimoError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise MimoError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C mimo_unet_b200/csrc`. There is no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().mimo_last_error().decode("utf-8", "replace")
        raise MimoError(f"{what or 'mimo_b200 call'} failed (status {rc}): {msg}")
