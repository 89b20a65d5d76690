"""ctypes binding of libskydome_b200.so (include/skydome_b200.h).  There is no fallback: if the library is missing the
import raises, and every op requires CUDA tensors."""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libskydome_b200.so")

OK = 0
ERR_INVALID, ERR_UNDEFINED_COORDS, ERR_CUDA, ERR_UNSUPPORTED, ERR_EVEN_KERNEL, ERR_NAN_OFFSET = -1, -2, -3, -4, -5, -6
EPI_NONE, EPI_LEAKY_RELU, EPI_RESIDUAL, EPI_RELU, EPI_LOG_DECOMPRESS, EPI_SUN_BLEND, EPI_MASK, EPI_FORCE_DIRECT = 0, 1, 2, 4, 8, 16, 32, 256
EPI_FORCE_BAND = 512
EPI_NO_PAIR = 1024
MATH_TF32, MATH_3XTF32 = 0, 1

_vp, _i, _f, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t

# name -> (restype, argtypes): every symbol include/skydome_b200.h declares
SIGNATURES = {
    "sky_version": (_i, []),
    "sky_last_error": (ctypes.c_char_p, []),
    "sky_launch_count": (ctypes.c_long, []),
    "sky_da_offsets_host": (_i, [_i, _i, _i, _i, _i, _vp]),
    "sky_da_offsets_device": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "sky_da_sample_debug": (_i, [_i, _i, _i, _vp] + [_vp] * 8 + [_vp]),
    "sky_da_packed_weight_bytes": (_sz, [_i, _i, _i, _i]),
    "sky_da_pack_weights": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "sky_da_conv2d_fwd": (_i, [_vp] * 8 + [_i, _i, _i, _i, _i, _i, _i, _f, _i, _vp]),
    "sky_da_strip_weight_bytes": (_sz, [_vp, _i, _i, _i, _i, _i, _i]),
    "sky_da_strip_pack_weights": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "sky_da_conv2d_fwd_strip": (_i, [_vp] * 7 + [_i, _i, _i, _i, _i, _i, _i, _f, _i, _vp]),
    "sky_da_strip_plan_info": (_i, [_vp, _i, _i, _i, _i, _vp]),
    "sky_da_strip_wgrad_plan_info": (_i, [_vp, _i, _i, _i, _i, _i, _vp]),
    "sky_conv_strip_wgrad_plan_info": (_i, [_i, _i, _i, _i, _i, _i, _vp]),
    "sky_conv_strip_wgrad_plan_export": (_i, [_i, _i, _i, _i, _i, _i] + [_vp] * 5),
    "sky_da_strip_wgrad_plan_export": (_i, [_vp, _i, _i, _i, _i, _i] + [_vp] * 7),
    "sky_da_strip_plan_export": (_i, [_vp, _i, _i, _i, _i] + [_vp] * 5),
    "sky_da_strip_weight_bytes_t": (_sz, [_vp, _i, _i, _i, _i, _i, _i]),
    "sky_da_strip_pack_weights_t": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "sky_da_conv2d_bwd_data_strip": (_i, [_vp] * 5 + [_i] * 8 + [_f, _i, _vp]),
    "sky_conv_strip_plan_info": (_i, [_i] * 9 + [_vp]),
    "sky_conv_strip_plan_export": (_i, [_i] * 9 + [_vp] * 3),
    "sky_conv2d_fwd": (_i, [_vp] * 6 + [_i] * 8 + [_f, _i, _vp]),
    "sky_conv2d_smallc_fwd": (_i, [_vp] * 5 + [_i] * 7 + [_f, _vp]),
    "sky_da_conv2d_smallc_fwd": (_i, [_vp] * 7 + [_i] * 7 + [_f, _vp]),
    "sky_da_conv2d_fwd_simt": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "sky_da_conv2d_bwd_data": (_i, [_vp] * 4 + [_i] * 7 + [_vp]),
    "sky_da_conv2d_bwd_filter": (_i, [_vp] * 5 + [_i] * 6 + [_vp]),
    "sky_da_conv2d_bwd_filter_strip": (_i, [_vp] * 5 + [_i] * 7 + [_vp]),
    "sky_debug_wgrad_trace": (_i, [_vp]),
    "sky_wgrad_schedule_info": (_i, [_i, _i, _i, _i, _i, _vp]),
    "sky_instnorm_apply": (_i, [_vp] * 6 + [_i, _i, _i, _i, _f, _i, _f, _vp]),
    "sky_instnorm_bwd": (_i, [_vp] * 10 + [_i, _i, _i, _i, _f, _f, _vp]),
    "sky_mse_loss": (_i, [_vp] * 4 + [ctypes.c_long, _vp]),
    "sky_rmsprop_step": (_i, [_vp] * 3 + [ctypes.c_long, _f, _f, _f, _f, _vp]),
    "sky_maxpool2x2_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "sky_dense_fwd": (_i, [_vp] * 4 + [_i, _i, _i, _i, _vp]),
    "sky_softmax_rows": (_i, [_vp, _vp, _i, _i, _vp]),
    "sky_debug_band_trace": (_i, [_vp]),
    "sky_debug_strip_trace": (_i, [_vp]),
    "sky_resize_bilinear_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "sky_conv2d_fwd_blend": (_i, [_vp] * 6 + [_f] + [_i] * 6 + [_f, _i, _vp]),
    "sky_softmax_max_bwd": (_i, [_vp] * 4 + [_i, _i, _vp]),
    "sky_softmax_pick_bwd": (_i, [_vp] * 5 + [_i, _i, _vp]),
    "sky_argmax_rows": (_i, [_vp, _vp, _i, _i, _vp]),
    "sky_blend_split": (_i, [_vp, _vp, _f] + [_vp] * 5 + [ctypes.c_long, _vp]),
    "sky_transpose": (_i, [_vp, _vp, _i, _i, _vp]),
    "sky_dense_bwd_data": (_i, [_vp] * 4 + [_i, _i, _i, _vp]),
    "sky_dense_bwd_data_nt": (_i, [_vp] * 4 + [_i, _i, _i, _vp]),
    "sky_maxpool2x2_bwd": (_i, [_vp] * 3 + [_i] * 4 + [_vp]),
    "sky_gradcam": (_i, [_vp] * 4 + [_i] * 4 + [_vp]),
    "sky_sunrad_input": (_i, [_vp] * 5 + [_i] * 8 + [_vp]),
    "sky_bn_fold": (_i, [_vp] * 5 + [_f, _vp, _vp, ctypes.c_long, _i, _vp]),
    "sky_max_nonneg": (_i, [_vp, _vp, ctypes.c_long, _vp]),
    "sky_ldr_synth": (_i, [_vp] * 9 + [_i] * 5 + [_vp]),
    "sky_hdr_log_codec": (_i, [_vp, _vp, ctypes.c_long, _i, _vp]),
    "sky_loss_reduce": (_i, [_i, _vp, _vp, ctypes.c_long, _vp, _vp]),
    "sky_kl_divergence": (_i, [_vp, _vp, ctypes.c_long, _vp, _vp]),
    "sky_dog_base": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "sky_dog_l1": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "sky_adam_step": (_i, [_vp] * 4 + [ctypes.c_long, _f, _f, _f, _f, ctypes.c_long, _f, _vp]),
    "sky_kl_divergence_bwd": (_i, [_vp, _vp, _vp, ctypes.c_long, _f, _i, _vp]),
    "sky_dog_l1_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "sky_dog_base_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "sky_softmax_bwd_rows": (_i, [_vp] * 4 + [_i, _i, _vp]),
    "sky_dense_bwd_filter": (_i, [_vp] * 4 + [_i, _i, _i, _vp]),
    "sky_da_conv2d_smallc_bwd_filter": (_i, [_vp] * 5 + [_i] * 6 + [_vp]),
    "sky_rgbe_encode": (_i, [_vp, _vp, ctypes.c_long, _i, _vp]),
    "sky_concat2_pad": (_i, [_vp, _i, _vp, _i, _vp, _i, ctypes.c_long, _vp]),
    "sky_vgg_preprocess": (_i, [_vp, _vp, ctypes.c_long, _f, _f, _f, _vp]),
    "sky_sun_radiance": (_i, [_vp] * 5 + [_i, _i, _f, _vp]),
    "sky_conv2d_transpose_weights": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "sky_conv2d_pack_weights_t": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "sky_conv2d_bwd_data": (_i, [_vp] * 4 + [_i] * 8 + [_f, _i, _vp]),
    "sky_conv2d_bwd_filter": (_i, [_vp] * 5 + [_i] * 9 + [_vp]),
    "sky_bn_train_stats": (_i, [_vp] * 5 + [_i, _i, _i, _i, _f, _vp]),
    "sky_bn_train_apply": (_i, [_vp] * 5 + [_i, _i, _i, _i, _f, _i, _f, _vp]),
    "sky_bn_train_bwd": (_i, [_vp] * 9 + [_i, _i, _i, _i, _f, _f, _vp]),
    "sky_resize_bilinear_bwd": (_i, [_vp, _vp] + [_i] * 7 + [_vp]),
    "sky_sun_radiance_bwd": (_i, [_vp] * 7 + [_i, _i, _f, _vp]),
    "sky_maxnorm_bwd": (_i, [_vp] * 5 + [ctypes.c_long, _i, _vp]),
    "sky_sunrad_heads_bwd": (_i, [_vp] * 6 + [_i, _i, _vp]),
    "sky_train_tail_fwd": (_i, [_vp] * 5 + [_f, _f] + [_vp] * 6 + [ctypes.c_long, _vp]),
    "sky_train_tail_bwd": (_i, [_vp] * 10 + [_f, _f, _f] + [_vp] * 3 + [ctypes.c_long, _vp]),
    "sky_lsgan_bwd": (_i, [_vp] * 3 + [_i] * 7 + [_f, _f, _vp]),
    "sky_l1_bwd": (_i, [_vp] * 4 + [ctypes.c_long, _f, _i, _vp]),
    "sky_maxpool2x2_bwd_relu": (_i, [_vp] * 3 + [_i] * 5 + [_vp]),
    "sky_zero": (_i, [_vp, _sz, _vp]),
}


class SkydomeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"skydome_b200 error {code}: {msg}")
        self.code = code


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library has not been built (run `python __graft_entry__.py` or "
            f"`python {os.path.join(HERE, 'build.py')}`).  There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == the .so does not export what the header declares
        fn.restype, fn.argtypes = res, args
    return lib


# kernel launches behind one call of each entry point (memsets not counted); bench.py's gpu_launches is derived from the calls a step
# makes.  Entry points not listed launch one kernel.
LAUNCHES_PER_CALL = {"sky_bn_train_stats": 2, "sky_bn_train_bwd": 2, "sky_conv2d_bwd_filter": 2, "sky_dense_fwd": 2, "sky_dense_bwd_data": 1, "sky_instnorm_bwd": 2, "sky_gradcam": 2, "sky_da_conv2d_bwd_filter": 2, "sky_da_conv2d_bwd_filter_strip": 2,
                     "sky_da_offsets_host": 0, "sky_zero": 0, "sky_da_packed_weight_bytes": 0, "sky_da_strip_weight_bytes": 0, "sky_da_strip_weight_bytes_t": 0, "sky_da_strip_plan_info": 0, "sky_da_strip_plan_export": 0, "sky_da_strip_wgrad_plan_info": 0, "sky_da_strip_wgrad_plan_export": 0, "sky_conv_strip_wgrad_plan_info": 0, "sky_conv_strip_wgrad_plan_export": 0,
                     "sky_conv_strip_plan_info": 0, "sky_conv_strip_plan_export": 0, "sky_last_error": 0, "sky_version": 0,
                     "sky_debug_band_trace": 0, "sky_debug_strip_trace": 0, "sky_debug_wgrad_trace": 0, "sky_wgrad_schedule_info": 0}


_UNTRACED = ("sky_launch_count", "sky_last_error", "sky_version", "sky_da_packed_weight_bytes", "sky_da_offsets_host", "sky_da_strip_weight_bytes", "sky_da_strip_weight_bytes_t",
             "sky_da_strip_plan_info", "sky_da_strip_plan_export", "sky_da_strip_wgrad_plan_info", "sky_da_strip_wgrad_plan_export", "sky_conv_strip_wgrad_plan_info", "sky_conv_strip_wgrad_plan_export", "sky_conv_strip_plan_info", "sky_conv_strip_plan_export")


class _Lib:
    """Attribute access to the C ABI with an optional per-entry-point call counter (`counts`: dict or None) and an optional device
    timeline (`trace`: list or None — every call is bracketed by CUDA events on the current stream; bench.py derives the kernel shares
    of a step and the roofline of its largest one from it)."""

    def __init__(self, cdll):
        self._cdll = cdll
        self.counts = None
        self.trace = None

    def __getattr__(self, name):
        fn = getattr(self._cdll, name)

        def call(*args):
            if self.counts is not None:
                self.counts[name] = self.counts.get(name, 0) + 1
            if self.trace is not None and name not in _UNTRACED:
                import torch
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                rc = fn(*args)
                e.record()
                self.trace.append((name, args, s, e))
                return rc
            return fn(*args)
        call.__name__ = name
        setattr(self, name, call)       # resolved once; later lookups bypass __getattr__
        return call

    def launches(self):
        return sum(n * LAUNCHES_PER_CALL.get(k, 1) for k, n in (self.counts or {}).items())


LIB = _Lib(load())


def check(rc):
    if rc != OK:
        msg = LIB.sky_last_error().decode("utf-8", "replace")
        if rc == ERR_EVEN_KERNEL:
            raise AssertionError(msg)                        # reference: assert at distortion_aware_ops.py:188
        if rc == ERR_UNDEFINED_COORDS:
            raise Exception("undefined coordinates")         # reference: distortion_aware_ops.py:252
        if rc == ERR_INVALID:
            raise ValueError(msg)
        raise SkydomeError(rc, msg)
    return rc
