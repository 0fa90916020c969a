"""ctypes binding of libnefnet_b200.so (include/nefnet_b200.h).

There is no fallback: if the shared library is missing, or the device is not sm_100, every entry point
raises.  PyTorch is only used to own device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("NEFNET_B200_LIB") or os.path.join(_PKG, "lib", "libnefnet_b200.so")   # override: A/B builds (tools/)

PHASE_TRAIN, PHASE_TEST, PHASE_GEN = 0, 1, 2
HALO = 3
GUARD_ROWS = 528


class NefConvTerm(C.Structure):
    _fields_ = [("x", C.c_void_p), ("x_cstride", C.c_int64), ("x_c4_off", C.c_int32), ("x_c4_gstride", C.c_int32),
                ("cin_g", C.c_int32), ("taps", C.c_int32), ("tap_off", C.c_int32), ("x_f16", C.c_int32),
                ("w", C.c_void_p)]


class NefConvDesc(C.Structure):
    _fields_ = [("n_terms", C.c_int32), ("groups", C.c_int32), ("N", C.c_int32), ("round_tf32", C.c_int32),
                ("term", NefConvTerm * 3), ("rows", C.c_int64), ("Lp", C.c_int32), ("L", C.c_int32),
                ("y", C.c_void_p), ("y_cstride", C.c_int64), ("y_c4_off", C.c_int32), ("y_c4_gstride", C.c_int32),
                ("y_Lp", C.c_int32), ("y_lmul", C.c_int32), ("y_ladd", C.c_int32), ("relu", C.c_int32),
                ("bias", C.c_void_p), ("res", C.c_void_p), ("res_cstride", C.c_int64), ("res_c4_off", C.c_int32),
                ("res_c4_gstride", C.c_int32), ("drop_p", C.c_float), ("mask_mode", C.c_int32),
                ("drop_seed", C.c_uint64), ("bscale", C.c_void_p), ("bscale_grad", C.c_void_p), ("mask", C.c_void_p),
                ("mask_cstride", C.c_int64), ("mask_c4_off", C.c_int32), ("mask_c4_gstride", C.c_int32),
                ("mask_scale", C.c_float), ("reserved2", C.c_int32), ("stat_sum", C.c_void_p),
                ("stat_sq", C.c_void_p), ("out_bits", C.c_void_p), ("mask_bits", C.c_void_p), ("y16", C.c_void_p),
                ("acc_scale", C.c_void_p), ("y16_scale", C.c_void_p), ("res16", C.c_void_p), ("res16_scale", C.c_void_p)]


class NefWgradDesc(C.Structure):
    _fields_ = [("dy", C.c_void_p), ("dy_cstride", C.c_int64), ("dy_c4_off", C.c_int32), ("dy_c4_gstride", C.c_int32),
                ("x", C.c_void_p), ("x_cstride", C.c_int64), ("x_c4_off", C.c_int32), ("x_c4_gstride", C.c_int32),
                ("cout_g", C.c_int32), ("cin_g", C.c_int32), ("groups", C.c_int32), ("taps", C.c_int32),
                ("tap_off", C.c_int32), ("wg_mod", C.c_int32), ("rows", C.c_int64), ("dw", C.c_void_p),
                ("sg", C.c_int64), ("sm", C.c_int64), ("sn", C.c_int64), ("st", C.c_int64), ("db", C.c_void_p)]


class NefForwardArgs(C.Structure):
    _fields_ = [("params", C.POINTER(C.c_void_p)), ("x", C.c_void_p), ("input_thetas", C.c_void_p),
                ("query_theta", C.c_void_p), ("rois", C.c_void_p), ("rest_theta", C.c_void_p), ("phase", C.c_int32),
                ("bn_training", C.c_int32), ("lead_choice_z1", C.c_int32), ("lead_choice_z2", C.c_int32),
                ("drop_p", C.c_float), ("save_for_backward", C.c_int32), ("drop_seed", C.c_uint64),
                ("out", C.c_void_p), ("out_p", C.c_void_p), ("out_l", C.c_void_p), ("rest_out", C.c_void_p)]


class NefBackwardArgs(C.Structure):
    _fields_ = [("params", C.POINTER(C.c_void_p)), ("grads", C.POINTER(C.c_void_p)), ("dout", C.c_void_p),
                ("dout_p", C.c_void_p), ("dout_l", C.c_void_p), ("ev_late_params_done", C.c_void_p)]


# name -> (restype, argtypes); every symbol include/nefnet_b200.h declares
SIGNATURES = {
    "nef_version": (C.c_int, []),
    "nef_last_error": (C.c_char_p, []),
    "nef_init": (C.c_int, [C.c_int]),
    "nef_set_conv_impl": (C.c_int, [C.c_int]),
    "nef_get_conv_impl": (C.c_int, []),
    "nef_set_exact_fp32": (C.c_int, [C.c_int]),
    "nef_launch_count": (C.c_int64, []),
    "nef_struct_size": (C.c_size_t, [C.c_int]),
    "nef_tc_dispatch_stats": (C.c_int, [C.POINTER(C.c_int64), C.c_int]),
    "nef_tc_dispatch_reset": (None, []),
    "nef_tc_set_persist_min": (C.c_int, [C.c_int]),
    "nef_param_count": (C.c_int, [C.c_int]),
    "nef_param_name": (C.c_char_p, [C.c_int, C.c_int]),
    "nef_param_numel": (C.c_int64, [C.c_int, C.c_int]),
    "nef_cbl4_rows": (C.c_int64, [C.c_int, C.c_int]),
    "nef_cbl4_floats": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "nef_ncl_to_cbl4": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nef_cbl4_to_ncl": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nef_ncl_to_h8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "nef_h8_to_ncl": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nef_pack_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                   C.c_int64, C.c_int64, C.c_int, C.c_void_p]),
    "nef_gconv_fwd": (C.c_int, [C.POINTER(NefConvDesc), C.c_void_p]),
    "nef_gconv_wgrad": (C.c_int, [C.POINTER(NefWgradDesc), C.c_void_p]),
    "nef_gconv_wgrad_f16": (C.c_int, [C.POINTER(NefWgradDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nef_plan_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "nef_plan_create_v": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "nef_param_count_v": (C.c_int, [C.c_int, C.c_int]),
    "nef_param_name_v": (C.c_char_p, [C.c_int, C.c_int, C.c_int]),
    "nef_param_numel_v": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "nef_plan_destroy": (None, [C.c_void_p]),
    "nef_plan_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "nef_plan_bind": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "nef_plan_tensor_info": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "nef_plan_export": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p]),
    "nef_plan_poison": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p]),
    "nef_forward": (C.c_int, [C.c_void_p, C.POINTER(NefForwardArgs), C.c_void_p]),
    "nef_backward": (C.c_int, [C.c_void_p, C.POINTER(NefBackwardArgs), C.c_void_p]),
    "nef_gen_ecg": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_int, C.c_void_p, C.c_void_p]),
    "nef_roi_check": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nef_loss_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                               C.POINTER(C.c_float), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nef_loss_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                               C.POINTER(C.c_float), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p]),
    "nef_pair_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nef_sgd_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_float,
                               C.c_void_p]),
    "nef_set_dec1_terms": (C.c_int, [C.c_int]),
    "nef_set_fwd_f16": (C.c_int, [C.c_int]),
    "nef_set_dec_f16": (C.c_int, [C.c_int]),
    "nef_psnr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                           C.c_void_p, C.c_void_p]),
    "nef_prepare_scratch_bytes": (C.c_size_t, [C.c_int]),
    "nef_prepare_segments": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                       C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "nef_stem_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nef_stem_tc_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nef_stem_tc_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nef_stem_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                               C.c_void_p]),
    "nef_angular_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "nef_angular_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None
_lock = threading.Lock()
_inited_devices = set()


def load():
    """dlopen the library (no CUDA call is made) and attach the signatures."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    "libnefnet_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                    "or `python electrocardio-panorama_b200/build.py`; there is no fallback path." % LIB_PATH)
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError if the symbol is missing
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (what or "nefnet_b200", rc, load().nef_last_error().decode()))


def init(device_index: int):
    lib = load()
    if device_index not in _inited_devices:
        check(lib.nef_init(int(device_index)), "nef_init")
        _inited_devices.add(device_index)
        if os.environ.get("NEF_FWD_F16"):  # measurement hook, see nef_set_fwd_f16
            check(lib.nef_set_fwd_f16(int(os.environ["NEF_FWD_F16"])), "nef_set_fwd_f16")
        terms = os.environ.get("NEF_DEC1_TERMS")  # measurement hook, see nef_set_dec1_terms
        if terms:
            check(lib.nef_set_dec1_terms(int(terms)), "nef_set_dec1_terms")
    return lib


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def guard(device):
    """Context manager: make `device` (a torch.device or a CUDA tensor) the current CUDA device for the native calls inside,
    so that stream_ptr() is that device's current stream and nef_init binds the right device -- a module on cuda:1 called
    while cuda:0 is current launches on cuda:1."""
    import torch
    if isinstance(device, torch.Tensor):
        device = device.device
    if device.type != "cuda":
        raise RuntimeError("nefnet_b200: CUDA tensors required (got %s); there is no CPU path" % device)
    return torch.cuda.device(device)


def device_index(device):
    import torch
    if isinstance(device, torch.Tensor):
        device = device.device
    return device.index if device.index is not None else torch.cuda.current_device()


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def dispatch_stats():
    """(counts[8], set of specialised epilogue codes the persistent kernel ran with) since the last reset; test hook."""
    buf = (C.c_int64 * 72)()
    n = load().nef_tc_dispatch_stats(buf, 72)
    return list(buf[:8]), set(int(buf[8 + i]) for i in range(n))


def param_names(G: int):
    lib = load()
    return [lib.nef_param_name(G, i).decode() for i in range(lib.nef_param_count(G))]
