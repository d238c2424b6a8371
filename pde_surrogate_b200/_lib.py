"""ctypes binding of the C-ABI declared in include/pdes_b200.h.

The product path has no CPU fallback: if the shared library is missing or a call fails, a
RuntimeError carrying pdes_last_error() is raised.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t,
                    c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libpdes_b200.so")
_lib = None


class DensenetConfig(Structure):
    _fields_ = [("in_channels", c_int32), ("out_channels", c_int32), ("imsize", c_int32),
                ("n_blocks", c_int32), ("blocks", c_int32 * 15), ("growth_rate", c_int32),
                ("init_features", c_int32), ("max_batch", c_int32), ("arch", c_int32), ("dropout", c_int32), ("upsample", c_int32), ("bottleneck", c_int32)]


class ConvDesc(Structure):
    _fields_ = [(n, c_int32) for n in ("B", "Hin", "Win", "Cin", "ld_in", "Hout", "Wout", "Cout", "ld_out",
                                       "c_off_out", "KH", "KW", "stride", "pad", "upsample", "bn_relu",
                                       "out_nchw")]


# name -> (restype, argtypes); mirrors include/pdes_b200.h one to one
SIGNATURES = {
    "pdes_last_error": (c_char_p, []),
    "pdes_abi_version": (c_int, []),
    "pdes_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int)]),
    "pdes_sobel_grad": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pdes_darcy_loss_workspace_bytes": (c_size_t, []),
    "pdes_darcy_loss_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "pdes_darcy_loss_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pdes_darcy_loss_nl_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p,
                                       c_void_p, c_void_p]),
    "pdes_darcy_loss_nl_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float,
                                       c_void_p, c_void_p]),
    "pdes_darcy_loss_set_impl": (c_int, [c_int]),
    "pdes_densenet_create": (c_int, [POINTER(DensenetConfig), POINTER(c_void_p)]),
    "pdes_densenet_destroy": (None, [c_void_p]),
    "pdes_densenet_num_params": (c_int, [c_void_p]),
    "pdes_densenet_param_floats": (c_int64, [c_void_p]),
    "pdes_densenet_param_info": (c_int, [c_void_p, c_int, c_char_p, c_size_t, POINTER(c_int64), POINTER(c_int32),
                                         POINTER(c_int64), POINTER(c_int32)]),
    "pdes_densenet_num_bn": (c_int, [c_void_p]),
    "pdes_densenet_running_floats": (c_int64, [c_void_p]),
    "pdes_densenet_bn_info": (c_int, [c_void_p, c_int, c_char_p, c_size_t, POINTER(c_int64), POINTER(c_int64),
                                      POINTER(c_int32)]),
    "pdes_densenet_output_size": (c_int, [c_void_p]),
    "pdes_densenet_dropout_sites": (c_int, [c_void_p, POINTER(c_int32), c_int]),
    "pdes_densenet_set_dropout": (c_int, [c_void_p, c_void_p]),
    "pdes_densenet_workspace_bytes": (c_size_t, [c_void_p]),
    "pdes_densenet_bind": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t]),
    "pdes_densenet_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "pdes_densenet_backward": (c_int, [c_void_p, c_void_p, c_void_p]),
    "pdes_densenet_backward_dx": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "pdes_densenet_flops": (c_double, [c_void_p, c_int, c_int]),
    "pdes_densenet_last_launches": (c_int, [c_void_p]),
    "pdes_densenet_set_conv_impl": (c_int, [c_void_p, c_int]),
    "pdes_densenet_set_timing": (c_int, [c_void_p, c_int]),
    "pdes_densenet_timing_report": (c_int, [c_void_p]),
    "pdes_densenet_timing_read": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_size_t]),
    "pdes_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float,
                               c_float, c_float, c_float, c_int64, c_void_p]),
    "pdes_adam_step_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "pdes_adam_hyper": (c_int, [c_void_p, c_float, c_float, c_float, c_float, c_float, c_float, c_int64]),
    "pdes_conv2d_fwd": (c_int, [POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_int, c_void_p]),
    "pdes_conv2d_dgrad": (c_int, [POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "pdes_conv2d_set_precision": (c_int, [c_int]),
    "pdes_conv_tc_plan": (c_int, [c_int, c_int, c_int, POINTER(c_int64)]),
    "pdes_conv2d_wgrad": (c_int, [POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                  c_void_p]),
}


def lib():
    """Load libpdes_b200.so (built by pde_surrogate_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            "pde_surrogate_b200: %s is missing. Build it with `python -m pde_surrogate_b200.build` "
            "(there is no CPU fallback for the CUDA path)." % SO_PATH)
    L = ctypes.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc, what=""):
    if rc != 0:
        msg = lib().pdes_last_error()
        raise RuntimeError("pdes_b200 %s failed (code %d): %s" % (what, rc, msg.decode() if msg else "?"))


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)
