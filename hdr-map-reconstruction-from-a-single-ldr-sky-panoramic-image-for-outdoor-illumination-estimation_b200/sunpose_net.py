"""Host-side mirror of the reference's sunpose_net.py: ``sunposeLayer(filter_out, k_h, k_w, strides, dilation_rate)`` and
``model(im_height, im_width, da_kernel_size, dilation_rate)`` with ``sunposeEstimation(x) -> (softmax, [act1, act2, act3])``.

`distortion_aware=True` (default) wires the convolutions the way the commented lines sunpose_net.py:11,16 do —
``distortion_aware_ops.conv2d(filter_out, kernel_size=k_h, strides=strides, dilation_rate=dilation_rate)`` — which is the
north-star path; `distortion_aware=False` gives the plain ``ops.conv2d`` wiring that is live in the committed reference."""
from __future__ import annotations

import math

import torch

from . import _lib, ops
from ._lib import LIB, check
from .distortion_aware_ops import _initializer, _require_cuda, _stream, conv2d as da_conv2d
from .generator import InstanceNormalization


def maxpool2d(x):
    """ops.maxpool2d(kernel_size=2).call (ops.py:299-300): tf.nn.max_pool 2x2 / 2 SAME."""
    x = _require_cuda(x, "x")
    B, h, w, C = x.shape
    y = torch.empty((B, (h + 1) // 2, (w + 1) // 2, C), dtype=torch.float32, device=x.device)
    check(LIB.sky_maxpool2x2_fwd(x.data_ptr(), y.data_ptr(), B, h, w, C, _stream()))
    return y


class Dense:
    """tf.keras.layers.Dense(units): kernel [in, units] glorot_uniform, bias zeros."""

    def __init__(self, units, device="cuda"):
        self.units = int(units)
        self.device = torch.device(device)
        self.kernel = None
        self.bias = None

    def build(self, input_shape):
        k = int(input_shape[-1])
        self.kernel = _initializer("glorot_uniform", (k, self.units), k, self.units, self.device)
        self.bias = torch.zeros(self.units, dtype=torch.float32, device=self.device)

    @property
    def trainable_variables(self):
        return [self.kernel, self.bias]

    def __call__(self, x, relu=False):
        x = _require_cuda(x, "x")
        if self.kernel is None:
            self.build(tuple(x.shape))
        B, K = x.shape
        y = torch.empty((B, self.units), dtype=torch.float32, device=x.device)
        check(LIB.sky_dense_fwd(x.data_ptr(), self.kernel.data_ptr(), self.bias.data_ptr(), y.data_ptr(), B, K, self.units,
                                int(bool(relu)), _stream()))
        return y


def softmax(x):
    x = _require_cuda(x, "x")
    rows, n = x.shape
    y = torch.empty_like(x)
    check(LIB.sky_softmax_rows(x.data_ptr(), y.data_ptr(), rows, n, _stream()))
    return y


class sunposeLayer:
    def __init__(self, filter_out, k_h=3, k_w=3, strides=1, dilation_rate=1, *, distortion_aware=True, math_mode=None,
                 device="cuda"):
        def conv():
            if distortion_aware:                               # sunpose_net.py:11,16
                return da_conv2d(filter_out, kernel_size=k_h, strides=strides, dilation_rate=dilation_rate,
                                 math_mode=math_mode, device=device)
            return ops.conv2d(output_channels=filter_out, k_h=k_h, k_w=k_w, strides=strides, math_mode=math_mode,
                              device=device)                   # :10,15
        self.filter_out = filter_out
        self.conv1, self.conv2 = conv(), conv()
        self.norm1, self.norm2 = InstanceNormalization(device=device), InstanceNormalization(device=device)
        self._stats = None

    @property
    def trainable_variables(self):
        return (self.conv1.trainable_variables + self.norm1.trainable_variables + self.conv2.trainable_variables
                + self.norm2.trainable_variables)

    def call(self, x, training="training"):
        x = _require_cuda(x, "x")
        B = x.shape[0]
        if self._stats is None or self._stats.shape[1] != B:
            self._stats = torch.zeros(2, B, self.filter_out, 2, dtype=torch.float64, device=x.device)
        else:
            self._stats.zero_()
        conv1 = self.conv1(x, stats=self._stats[0])                                  # :21
        actv1 = self.norm1.apply(conv1, self._stats[0], leaky_slope=0.0)             # :22-23  relu == leaky_relu(0)
        conv2 = self.conv2(actv1, stats=self._stats[1])                              # :25
        return self.norm2.apply(conv2, self._stats[1], leaky_slope=0.0)              # :26-29

    __call__ = call


class model:
    def __init__(self, im_height=32, im_width=128, da_kernel_size=3, dilation_rate=1, *, distortion_aware=True,
                 math_mode=None, device="cuda"):
        self.fc_dim = int(im_height * im_width)                                      # sunpose_net.py:36
        kw = dict(distortion_aware=distortion_aware, math_mode=math_mode, device=device)
        self.sunlayer1 = sunposeLayer(32, k_h=7, k_w=7, **kw)                        # :39
        self.sunlayer2 = sunposeLayer(64, k_h=3, k_w=3, **kw)                        # :42
        self.sunlayer3 = sunposeLayer(128, k_h=3, k_w=3, **kw)                       # :45
        self.fc1 = Dense(self.fc_dim, device=device)                                 # :49
        self.fc2 = Dense(self.fc_dim, device=device)                                 # :51

    @property
    def trainable_variables(self):
        return (self.sunlayer1.trainable_variables + self.sunlayer2.trainable_variables + self.sunlayer3.trainable_variables
                + self.fc1.trainable_variables + self.fc2.trainable_variables)

    def sunposeEstimation(self, x, training="training"):
        sunlayer1 = self.sunlayer1(x, training)                                      # :55
        pool1_s = maxpool2d(sunlayer1)
        sunlayer2 = self.sunlayer2(pool1_s, training)
        pool2_s = maxpool2d(sunlayer2)
        sunlayer3 = self.sunlayer3(pool2_s, training)
        pool3_s = maxpool2d(sunlayer3)
        flat = pool3_s.reshape(pool3_s.shape[0], -1)                                 # Keras Flatten on NHWC: (h, w, c) order
        actv1_s = self.fc1(flat, relu=True)                                          # :65-66
        actv2_s = self.fc2(actv1_s, relu=True)                                       # :67-68
        sm = softmax(actv2_s)                                                        # :70
        return sm, [sunlayer1, sunlayer2, sunlayer3]
