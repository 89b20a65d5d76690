"""Host-side mirror of the plain layers of the reference's ops.py that the generator instantiates:

    ops.conv2d(output_channels, strides, k_h, k_w, padding="SAME")            ops.py:4-42   (tf.nn.conv2d + bias_add)
    ops.deconv2d(output_channels, output_imshape, k_h, k_w, method='resize')  ops.py:44-126 (tf.image.resize BILINEAR,
                                                                               then a stride-1 SAME conv; only the
                                                                               'resize' method is used by generator.py)

Variables keep the reference's names and shapes: ``w`` / ``kernel`` [k_h, k_w, C, F], ``biases`` [F].  Layers with
C % 32 == 0 run the identity-sampler mode of the band-staged tcgen05 kernel (stride 1 or 2), image-like inputs (C <= 4)
the fp32 small-C kernel, anything else the direct-gather kernel (all behind sky_conv2d_fwd / sky_conv2d_smallc_fwd).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib
from ._lib import LIB, check
from .distortion_aware_ops import _MATH, DEFAULT_MATH_MODE, _initializer, _ptr, _require_cuda, _stream, resize_bilinear


def _epilogue_flags(leaky_slope, residual, relu, log_decompress):
    flags = _lib.EPI_NONE
    if leaky_slope is not None:
        flags |= _lib.EPI_LEAKY_RELU
    if residual is not None:
        flags |= _lib.EPI_RESIDUAL
    if relu:
        flags |= _lib.EPI_RELU
    if log_decompress:
        flags |= _lib.EPI_LOG_DECOMPRESS
    return flags


class TransposedPack:
    """Cache of the flipped, transposed kernel of a plain conv layer — unpacked [k*k*F, C] (`scratch`) and packed for the tensor-core
    kernels — which turns the layer's data gradient into a forward pass over dy (sky_conv2d_bwd_data)."""

    def __init__(self, k, C, F, math_mode, device):
        self.k, self.C, self.F, self.mode = k, C, F, _MATH[math_mode]
        self.scratch = torch.empty(k * k * F * C, dtype=torch.float32, device=device)
        self.packed = torch.empty(LIB.sky_da_packed_weight_bytes(F, C, k, self.mode), dtype=torch.uint8, device=device)
        self.key = None

    def get(self, w):
        key = (w.data_ptr(), w._version)
        if self.key != key:
            if self.F <= 4:      # 3-filter layers also run on the unpacked transposed kernel (small-C kernel of conv_backward_data)
                check(LIB.sky_conv2d_transpose_weights(w.data_ptr(), self.scratch.data_ptr(), self.C, self.F, self.k, 1, _stream()))
            check(LIB.sky_conv2d_pack_weights_t(w.data_ptr(), self.packed.data_ptr(), self.C, self.F, self.k, self.mode, _stream()))
            self.key = key
        return self.packed


def conv_backward_data(tp: TransposedPack, w, in_shape, dy, stride=1, mask_src=None, slope=0.0, residual=None):
    """dx [in_shape] of a plain SAME conv with kernel `w` ([k,k,C,F]) for the upstream gradient dy; mask_src: the activation tensor
    that fed the conv (its LeakyReLU / ReLU gradient is applied in the epilogue); residual: a gradient to add (identity branch)."""
    B, h, w_, C = in_shape
    dy = _require_cuda(dy, "dy")
    dx = torch.empty(in_shape, dtype=torch.float32, device=dy.device)
    k, F = tp.k, tp.F
    packed = tp.get(w)
    if F <= 4 and C <= 32 and stride == 1 and k % 2 == 1 and k <= 11 and mask_src is None and residual is None:
        # 3-filter layers (conv1_f / conv1_u): the transposed conv has a 3-channel input -> the fp32 small-C kernel on the unpacked kernel
        check(LIB.sky_conv2d_smallc_fwd(dy.data_ptr(), tp.scratch.data_ptr(), None, dx.data_ptr(), None, B, h, w_, F, C, k, 0, 0.0, _stream()))
        return dx
    if mask_src is not None and residual is not None:
        raise ValueError("mask_src and residual share one epilogue slot")
    flags = _lib.EPI_MASK if mask_src is not None else (_lib.EPI_RESIDUAL if residual is not None else 0)
    aux = mask_src if mask_src is not None else residual
    check(LIB.sky_conv2d_bwd_data(dy.data_ptr(), packed.data_ptr(), dx.data_ptr(), _ptr(aux), B, h, w_, C, F, k, int(stride), flags,
                                  float(slope), tp.mode, _stream()))
    return dx


_ZERO_TABLES = {}


def conv_backward_filter(x, dy, k, stride, dk_out, db_out=None, accumulate=False, C_store=None):
    """dkernel [k*k*C_store, F] (+ dbias) of a plain SAME conv at input x for the upstream gradient dy."""
    x, dy = _require_cuda(x, "x"), _require_cuda(dy, "dy")
    B, h, w_, C = x.shape
    F = dy.shape[-1]
    if C <= 4:
        # image-like input (conv1_d): the small-C distortion-aware weight-gradient kernel through an all-zero offset table (every tap on
        # a grid point == the plain SAME conv for odd k, stride 1)
        if stride != 1 or k % 2 == 0 or accumulate:
            raise NotImplementedError("small-C weight gradient: stride 1, odd kernel, overwrite")
        key = (h, k, str(x.device))
        if key not in _ZERO_TABLES:
            _ZERO_TABLES[key] = torch.zeros((h, k * k, 2), dtype=torch.float32, device=x.device)
        check(LIB.sky_da_conv2d_smallc_bwd_filter(x.data_ptr(), dy.data_ptr(), _ZERO_TABLES[key].data_ptr(), dk_out.data_ptr(),
                                                  _ptr(db_out), B, h, w_, C, F, k, _stream()))
        return
    check(LIB.sky_conv2d_bwd_filter(x.data_ptr(), dy.data_ptr(), None, dk_out.data_ptr(), _ptr(db_out), B, h, w_, C, int(C_store or C), F, k,
                                    int(stride), int(bool(accumulate)), _stream()))


class _PlainConvCore:
    """Shared by conv2d and deconv2d: weight packing and the two launch routes."""

    def __init__(self, output_channels, k_h, k_w, strides, kernel_initializer, bias_initializer, math_mode, device):
        if k_h != k_w:
            raise ValueError("only square kernels are used by the reference model")
        self.output_channels = output_channels
        self.k_h, self.k_w = k_h, k_w
        self.stride = int(strides)
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer
        self.math_mode = math_mode or DEFAULT_MATH_MODE
        self.device = torch.device(device)
        self.built = False
        self._packed = None
        self._packed_key = None
        self._tpack = None

    def _build_weights(self, channels_in):
        k, F = self.k_h, self.output_channels
        fan_in, fan_out = k * k * channels_in, k * k * F          # Keras 4-D glorot: receptive field x channels
        w = _initializer(self.kernel_initializer, (k, k, channels_in, F), fan_in, fan_out, self.device)
        b = _initializer(self.bias_initializer, (F,), fan_in, fan_out, self.device)
        self._channels_in = channels_in
        self.built = True
        return w, b

    def _weight(self):
        raise NotImplementedError

    def _bias(self):
        raise NotImplementedError

    def _packed_weights(self):
        wt = self._weight()
        mode = _MATH[self.math_mode]
        key = (wt.data_ptr(), wt._version, mode)
        if self._packed_key != key:
            k, C, F = self.k_h, self._channels_in, self.output_channels
            nbytes = LIB.sky_da_packed_weight_bytes(C, F, k, mode)
            if self._packed is None or self._packed.numel() != nbytes:
                self._packed = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            check(LIB.sky_da_pack_weights(wt.data_ptr(), self._packed.data_ptr(), C, F, k, mode, _stream()))
            self._packed_key = key
        return self._packed

    def _invalidate(self):
        self._packed_key = None
        if self._tpack is not None:
            self._tpack.key = None

    def backward_data(self, x, dy, mask_src=None, slope=0.0, residual=None):
        """Gradient of the SAME conv (stride 1 or 2, any kernel size) w.r.t. its input `x` (only its shape is used): a forward pass
        over dy with the flipped, transposed kernel (csrc/conv_bwd.cu) — a gather, no atomics."""
        if self._tpack is None:
            self._tpack = TransposedPack(self.k_h, self._channels_in, self.output_channels, self.math_mode, self.device)
        return conv_backward_data(self._tpack, self._weight(), tuple(x.shape), dy, self.stride, mask_src, slope, residual)

    def backward_filter(self, x, dy, dk_out, db_out=None, accumulate=False):
        """Weight / bias gradient of the SAME conv at input x (for deconv2d: the resized input)."""
        conv_backward_filter(x, dy, self.k_h, self.stride, dk_out.view(-1, self.output_channels), db_out, accumulate)

    def _conv(self, x, leaky_slope=None, residual=None, relu=False, log_decompress=False, stats=None, blend=None):
        B, h, w, C = x.shape
        k, F, s = self.k_h, self.output_channels, self.stride
        if C != self._channels_in:
            raise ValueError("channel count differs from the shape the layer was built for")
        oh, ow = -(-h // s), -(-w // s)
        y = torch.empty((B, oh, ow, F), dtype=torch.float32, device=x.device)
        if residual is not None:
            residual = _require_cuda(residual, "residual")
        flags = _epilogue_flags(leaky_slope, residual, relu, log_decompress)
        slope = float(leaky_slope or 0.0)
        mode = _MATH[self.math_mode]
        if blend is not None:
            # conv1_u with the tail of inference.generator_in_step in its epilogue (sky_conv2d_fwd_blend)
            sky_gamma, threshold = blend
            sky_gamma = _require_cuda(sky_gamma, "sky_gamma")
            if F != 3 or s != 1 or stats is not None or tuple(sky_gamma.shape) != (B, oh, ow, 3):
                raise ValueError("blend epilogue: 3 filters, stride 1, sky prediction of the output's shape")
            check(LIB.sky_conv2d_fwd_blend(x.data_ptr(), self._packed_weights().data_ptr(), self._bias().data_ptr(), y.data_ptr(),
                                           _ptr(residual), sky_gamma.data_ptr(), float(threshold), B, h, w, C, k, flags, slope,
                                           mode, _stream()))
        elif (s == 1 and k % 2 == 1 and k <= 11 and C <= 4 and F <= 256 and residual is None and not log_decompress
              and not (relu and leaky_slope is not None)):
            # image-like input (conv1_d; VGG conv1_1 with its ReLU): fp32 CUDA-core kernel, the unpacked variable is read directly
            check(LIB.sky_conv2d_smallc_fwd(x.data_ptr(), self._weight().data_ptr(), self._bias().data_ptr(), y.data_ptr(),
                                            _ptr(stats), B, h, w, C, F, k, flags, slope, _stream()))
        else:
            check(LIB.sky_conv2d_fwd(x.data_ptr(), self._packed_weights().data_ptr(), self._bias().data_ptr(), y.data_ptr(),
                                     _ptr(residual), _ptr(stats), B, h, w, C, F, k, s, flags, slope, mode, _stream()))
        return y


class conv2d(_PlainConvCore):
    def __init__(self, output_channels="output_channels", strides="strides", k_h="k_h", k_w="k_w", padding="SAME",
                 kernel_initializer='glorot_uniform', bias_initializer='zeros', *, math_mode=None, device="cuda"):
        if padding != "SAME":
            raise ValueError("the generator only uses SAME padding (ops.py:10)")
        super().__init__(output_channels, k_h, k_w, strides, kernel_initializer, bias_initializer, math_mode, device)
        self.strides = [1, strides, strides, 1]                 # ops.py:17
        self.padding = padding
        self.w = None
        self.biases = None

    def build(self, input_shape):
        self.w, self.biases = self._build_weights(int(input_shape[-1]))      # ops.py:30-38

    def _weight(self):
        return self.w

    def _bias(self):
        return self.biases

    @property
    def trainable_variables(self):
        return [self.w, self.biases]

    def call(self, input, **epilogue):
        x = _require_cuda(input, "input")
        if not self.built:
            self.build(tuple(x.shape))
        return self._conv(x, **epilogue)                                      # ops.py:41-42

    __call__ = call


class deconv2d(_PlainConvCore):
    def __init__(self, output_channels="output_channels", output_imshape=[], k_h="k_h", k_w="k_w", strides=1,
                 padding="SAME", method='resize', kernel_initializer='glorot_uniform', bias_initializer='zeros', *,
                 math_mode=None, device="cuda"):
        if method != 'resize':
            raise NotImplementedError("generator.py only instantiates ops.deconv2d(method='resize') (generator.py:70-83)")
        super().__init__(output_channels, k_h, k_w, 1, kernel_initializer, bias_initializer, math_mode, device)
        self.output_imshape = [int(v) for v in output_imshape]
        self.method = method
        self.kernel = None
        self.biases = None

    def build(self, input_shape):
        self.kernel, self.biases = self._build_weights(int(input_shape[-1]))  # ops.py:98-110

    def _weight(self):
        return self.kernel

    def _bias(self):
        return self.biases

    @property
    def trainable_variables(self):
        return [self.kernel, self.biases]

    def call(self, input, **epilogue):
        x = _require_cuda(input, "input")
        if not self.built:
            self.build(tuple(x.shape))
        im_resized = resize_bilinear(x, self.output_imshape[0], self.output_imshape[1])   # ops.py:122
        self._last_resized = im_resized                                                    # the conv's input: its weight gradient needs it
        return self._conv(im_resized, **epilogue)                                          # ops.py:124

    __call__ = call
