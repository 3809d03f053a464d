"""ctypes binding of libb200voc.so (include/b200voc.h).  There is no CPU fallback: if the shared
library is missing or the device is not sm_100, every compute entry point raises."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libb200voc.so")
# development build (`make dev`, include/b200voc_dev.h): the product library plus experiment / trace exports;
# only the debug drivers under tests/ select it, with B200VOC_LIB=dev
DEV_LIB_PATH = os.path.join(os.path.dirname(_HERE), "libb200voc_dev.so")

OK, ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_STATE, ERR_OVERFLOW = 0, -1, -2, -3, -4, -5
FMT_FP16, FMT_BF16 = 0, 1
PLAN_FP16, PLAN_BF16, PLAN_MIXED = 0, 1, 2
PLANS = {"fp16": PLAN_FP16, "bf16": PLAN_BF16, "mixed": PLAN_MIXED}


class GenConfig(C.Structure):
    _fields_ = [
        ("channels", C.c_int32), ("cond_dim", C.c_int32), ("style_dim", C.c_int32), ("num_bands", C.c_int32),
        ("n_stages", C.c_int32), ("upsample_factors", C.c_int32 * 8),
        ("n_dilations", C.c_int32), ("res_dilations", C.c_int32 * 8),
        ("hidden_dim", C.c_int32), ("use_attention", C.c_int32), ("attn_window", C.c_int32),
        ("precision_plan", C.c_int32), ("reserved", C.c_int32 * 7),
    ]


class GenIO(C.Structure):
    """b200voc_gen_io (include/b200voc.h): wire formats either side of Generator.forward."""
    _fields_ = [("mel_time_major", C.c_int32), ("out_format", C.c_int32), ("valid_samples", C.c_void_p),
                ("reserved", C.c_int32 * 4)]


OUT_F32, OUT_PCM16 = 0, 1
_P, _I, _I64, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float
_SIGNATURES = {
    "b200voc_version": (C.c_int, []),
    "b200voc_last_error_string": (C.c_char_p, []),
    "b200voc_device_supported": (C.c_int, [_I]),
    "b200voc_gen_create": (C.c_int, [C.POINTER(GenConfig), C.POINTER(_P)]),
    "b200voc_gen_set_weight": (C.c_int, [_P, C.c_char_p, _P, _I64, _P]),
    "b200voc_gen_num_weights": (C.c_int, [_P]),
    "b200voc_gen_weight_name": (C.c_char_p, [_P, _I]),
    "b200voc_gen_weight_numel": (_I64, [_P, _I]),
    "b200voc_gen_finalize": (C.c_int, [_P]),
    "b200voc_gen_workspace_bytes": (_I64, [_P, _I, _I]),
    "b200voc_gen_forward": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _P, _P, _I64, C.c_char_p, _P, _P]),
    "b200voc_gen_forward_ex": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, C.POINTER(GenIO), _P, _P, _I64,
                                         C.c_char_p, _P, _P]),
    "b200voc_gst_scratch_bytes": (_I64, [_I, _I, _I]),
    "b200voc_gst_forward": (C.c_int, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _I64, _P, _P]),
    "b200voc_disc_conv_out_len": (C.c_int, [_I, _I, _I, _I]),
    "b200voc_disc_conv": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I64, _I64, _F, _P, _P, _P]),
    "b200voc_disc_conv_tc_supported": (C.c_int, [_I, _I, _I, _I, _I]),
    "b200voc_disc_conv_tc_workspace_bytes": (C.c_int64, [_I, _I, _I]),
    "b200voc_disc_split_weight_elems": (C.c_int64, [_I, _I, _I]),
    "b200voc_disc_pack_weight_split": (C.c_int, [_P, _I, _I, _I, _P, _P]),
    "b200voc_disc_conv_tc": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _P, _P, _I64, _P]),
    "b200voc_spectral_norm_weight": (C.c_int, [_P, _P, _P, _I, _I, _P, _P, _P]),
    "b200voc_spectral_norm_train": (C.c_int, [_P, _P, _P, _I, _I, _F, _P, _P, _P, _P]),
    "b200voc_avg_pool1d_k4s2p1": (C.c_int, [_P, _I64, _I, _P, _P]),
    "b200voc_disc_lrelu_bwd": (C.c_int, [_P, _P, _P, _P, _F, _I64, _P, _P]),
    "b200voc_disc_bias_grad": (C.c_int, [_P, _I, _I, _I64, _P, _P]),
    "b200voc_disc_conv_dgrad": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I64, _I64, _I, _P, _P]),
    "b200voc_disc_conv_wgrad_scratch_bytes": (C.c_int64, [_I, _I, _I, _I, _I, _I, _I, _I]),
    "b200voc_disc_conv_wgrad": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I64, _I64, _P, _P, _P]),
    "b200voc_avg_pool1d_k4s2p1_bwd": (C.c_int, [_P, _I64, _I, _P, _P]),
    "b200voc_spectral_norm_bwd_scratch_bytes": (C.c_int64, []),
    "b200voc_spectral_norm_bwd": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _P, _P, _P]),
    "b200voc_disc_conv_dgrad_tc_supported": (C.c_int, [_I, _I, _I, _I, _I, _I]),
    "b200voc_disc_flip_weight": (C.c_int, [_P, _I, _I, _I, _P, _P]),
    "b200voc_disc_conv_wgrad_tc_supported": (C.c_int, [_I, _I, _I, _I, _I, _I, _I, _I]),
    "b200voc_disc_conv_wgrad_tc_workspace_bytes": (C.c_int64, [_I, _I, _I, _I, _I, _I]),
    "b200voc_disc_conv_wgrad_tc": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _I64, _P]),
    "b200voc_gen_launch_count": (C.c_int, [_P]),
    "b200voc_gen_set_overflow_check": (C.c_int, [_P, _I]),
    "b200voc_gen_profile_enable": (C.c_int, [_P, _I]),
    "b200voc_gen_profile_count": (C.c_int, [_P]),
    "b200voc_gen_profile_name": (C.c_char_p, [_P, _I]),
    "b200voc_gen_profile_ms": (C.c_float, [_P, _I]),
    "b200voc_gen_profile_flops": (C.c_double, [_P, _I]),
    "b200voc_gen_profile_bytes": (C.c_double, [_P, _I]),
    "b200voc_gen_destroy": (C.c_int, [_P]),
    "b200voc_convt_packed_elems": (_I64, [_I, _I, _I]),
    "b200voc_pack_convt_weight": (C.c_int, [_P, _I, _I, _I, _I, _P, _P]),
    "b200voc_convt1d": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "b200voc_resblock_packed_elems": (_I64, [_I]),
    "b200voc_resblock_input_is_lrelu": (_I, [_I]),
    "b200voc_pack_resblock_weights": (C.c_int, [_P, _P, _I, _I, _P, _P]),
    "b200voc_resblock": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "b200voc_merge_packed_elems": (_I64, [_I]),
    "b200voc_pack_merge_weight": (C.c_int, [_P, _I, _I, _P, _P]),
    "b200voc_stage_fused": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "b200voc_stft_prepare": (C.c_int, [_I, _I, _I]),
    "b200voc_stft_mag": (C.c_int, [_P, _I, _I, _I, _I, _P, _P, _P]),
    "b200voc_stft_complex": (C.c_int, [_P, _I, _I, _I, _I, _P, _P]),
    "b200voc_stft_logmel": (C.c_int, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "b200voc_istft": (C.c_int, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "b200voc_stft_l1": (C.c_int, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "b200voc_stft_l1_backward_workspace_bytes": (_I64, [_I, _I, _I, _I]),
    "b200voc_stft_l1_backward": (C.c_int, [_P, _P, _I, _I, _I, _I, _P, _F, _P, _P, _P, _I64, _P]),
    "b200voc_stft_mag_backward_workspace_bytes": (C.c_int64, [_I, _I, _I, _I]),
    "b200voc_stft_mag_backward": (C.c_int, [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _I64, _P]),
}
EXPORTS = tuple(_SIGNATURES)
_DEV_SIGNATURES = {
    "b200voc_exp_rowshift": (C.c_int, [_P, _P, _P, _P]),
    "b200voc_exp_cta2": (C.c_int, [_P, _P, _I, _P, _P, _P]),
    "b200voc_debug_set_trace": (C.c_int, [_P]),
    "b200voc_exp_mma_rate": (C.c_int, [_I, _I, _I, _P, _P]),
}
DEV_EXPORTS = tuple(_DEV_SIGNATURES)

_lib: Optional[C.CDLL] = None


class B200VocError(RuntimeError):
    pass


class B200VocOverflowError(B200VocError):
    """a 16-bit activation overflowed its storage format (raised only when the overflow check is enabled)"""


def load() -> C.CDLL:
    """dlopen the in-tree library and bind every symbol the header declares."""
    global _lib
    if _lib is not None:
        return _lib
    sel = os.environ.get("B200VOC_LIB", "")
    dev = sel == "dev"
    path = DEV_LIB_PATH if dev else LIB_PATH
    if sel.endswith(".so"):        # A/B runs of two builds on the same box (tests/ab_step.py); never set in production
        path = os.path.abspath(sel)
    if not os.path.exists(path):
        raise B200VocError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU / PyTorch fallback for the b200voc hot path)")
    lib = C.CDLL(path)
    sigs = dict(_SIGNATURES, **_DEV_SIGNATURES) if dev else _SIGNATURES
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status == OK:
        return
    msg = load().b200voc_last_error_string().decode("utf-8", "replace")
    text = f"b200voc {what} failed ({status}): {msg}"
    if status in (ERR_BAD_ARG, ERR_UNSUPPORTED):
        raise ValueError(text)
    if status == ERR_OVERFLOW:
        raise B200VocOverflowError(text)
    raise B200VocError(text)


def ptr(t) -> int:
    """device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise B200VocError("b200voc has no CPU path: tensors must live on a CUDA (sm_100) device")
