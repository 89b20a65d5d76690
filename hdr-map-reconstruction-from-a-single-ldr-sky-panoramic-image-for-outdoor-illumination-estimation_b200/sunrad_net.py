"""Host-side mirror of the reference's sunrad_net.py: ``downsampling(filters, kernel_size, strides=2, apply_norm=True)`` and
``sunRadNet(epsilon=1e-5)`` with ``call(x, actv_map, training) -> (sun_rad, gamma_in, beta_in)`` (sunrad_net.py:8-71).

Inference mode (training=False, inference.py:103): Keras BatchNormalization uses its moving statistics, so each
conv -> BN -> LeakyReLU(0.3) block is ONE conv launch: the BN scale is folded into the (bias-free) kernel and its shift into a
bias by ``sky_bn_fold`` whenever a variable changes, and LeakyReLU rides in the conv epilogue.  d4 has 512 filters and runs as
two 256-filter slices inside ``sky_conv2d_fwd``.  The two Dense(1) heads are one [32768, 2] weight-streaming product, and the
Dirac-delta-like radiance function, its 30000 clamp, hdr_logCompression and the x3 tile are one elementwise kernel.
The batch-statistics mode (training=True, train.py) is not built yet.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib
from ._lib import LIB, check
from .distortion_aware_ops import _MATH, DEFAULT_MATH_MODE, _ptr, _require_cuda, _stream


class downsampling:
    def __init__(self, filters, kernel_size, strides=2, apply_norm=True, *, math_mode=None, device="cuda", channel_pad=None):
        self.filters, self.kernel_size, self.strides, self.apply_norm = int(filters), int(kernel_size), int(strides), bool(apply_norm)
        self.math_mode = math_mode or DEFAULT_MATH_MODE
        self.device = torch.device(device)
        self.channel_pad = channel_pad          # pad input channels up to this count (6 -> 8 for d1) with zero kernel rows
        self.kernel = None                      # Keras Conv2D kernel [k, k, C, F], N(0, 0.02), use_bias=False (sunrad_net.py:12-14)
        self.gamma = self.beta = self.moving_mean = self.moving_variance = None     # BatchNormalization variables
        self.epsilon = 1e-3                     # Keras BatchNormalization default
        self.alpha = 0.3                        # Keras LeakyReLU default (sunrad_net.py:18)
        self._folded = self._fold_bias = self._packed = None
        self._key = None

    def build(self, input_shape):
        C, k, F = int(input_shape[-1]), self.kernel_size, self.filters
        self._channels = C
        self.kernel = torch.empty((k, k, C, F), dtype=torch.float32).normal_(0.0, 0.02).to(self.device)
        self.gamma = torch.ones(F, dtype=torch.float32, device=self.device)
        self.beta = torch.zeros(F, dtype=torch.float32, device=self.device)
        self.moving_mean = torch.zeros(F, dtype=torch.float32, device=self.device)
        self.moving_variance = torch.ones(F, dtype=torch.float32, device=self.device)

    @property
    def trainable_variables(self):
        return [self.kernel] + ([self.gamma, self.beta] if self.apply_norm else [])

    def _packed_weights(self):
        vs = (self.kernel, self.gamma, self.beta, self.moving_mean, self.moving_variance)
        key = tuple((v.data_ptr(), v._version) for v in vs)
        if key != self._key:
            k, C, F = self.kernel_size, self._channels, self.filters
            Cp = max(C, self.channel_pad or C)
            if self._folded is None:
                self._folded = torch.zeros((k, k, Cp, F), dtype=torch.float32, device=self.device)
                self._fold_bias = torch.zeros(F, dtype=torch.float32, device=self.device)
                nbytes = LIB.sky_da_packed_weight_bytes(Cp, F, k, _MATH[self.math_mode])
                self._packed = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            src = self.kernel
            if self.apply_norm:
                src = torch.empty_like(self.kernel)
                check(LIB.sky_bn_fold(self.kernel.data_ptr(), self.gamma.data_ptr(), self.beta.data_ptr(), self.moving_mean.data_ptr(),
                                      self.moving_variance.data_ptr(), self.epsilon, src.data_ptr(), self._fold_bias.data_ptr(),
                                      k * k * C, F, _stream()))
            self._folded[:, :, :C, :].copy_(src)
            check(LIB.sky_da_pack_weights(self._folded.data_ptr(), self._packed.data_ptr(), Cp, F, k, _MATH[self.math_mode], _stream()))
            self._key = key
        return self._packed

    def call(self, x, training="training"):
        if training is True or training == "training":
            raise NotImplementedError("sunRadNet with batch statistics (training=True) is not built; inference mode only")
        x = _require_cuda(x, "x")
        if self.kernel is None:
            self.build(tuple(x.shape))
        B, h, w, Cx = x.shape
        Cp = max(self._channels, self.channel_pad or self._channels)
        if Cx != Cp:
            raise ValueError(f"input has {Cx} channels, the layer expects {Cp} (channel-padded)")
        s = self.strides
        y = torch.empty((B, -(-h // s), -(-w // s), self.filters), dtype=torch.float32, device=x.device)
        packed = self._packed_weights()
        check(LIB.sky_conv2d_fwd(x.data_ptr(), packed.data_ptr(), self._fold_bias.data_ptr(), y.data_ptr(), None, None, B, h, w, Cp,
                                 self.filters, self.kernel_size, s, _lib.EPI_LEAKY_RELU, self.alpha, _MATH[self.math_mode], _stream()))
        return y

    __call__ = call


class sunRadNet:
    def __init__(self, epsilon=1e-5, pi=math.pi, *, math_mode=None, device="cuda"):
        self.epsilon = float(epsilon)
        self.device = torch.device(device)
        kw = dict(math_mode=math_mode, device=device)
        self.d1 = downsampling(64, 4, strides=2, apply_norm=False, channel_pad=8, **kw)    # sunrad_net.py:37
        self.d2 = downsampling(128, 4, strides=2, apply_norm=True, **kw)                   # :38
        self.d3 = downsampling(256, 4, strides=2, apply_norm=True, **kw)                   # :39
        self.d4 = downsampling(512, 4, strides=1, apply_norm=True, **kw)                   # :40
        self.gamma_kernel = self.gamma_bias = self.beta_kernel = self.beta_bias = None     # Dense(1) x 2 (:43-44)
        self._gb_kernel = self._gb_bias = None
        self._gb_key = None

    @property
    def trainable_variables(self):
        return (self.d1.trainable_variables + self.d2.trainable_variables + self.d3.trainable_variables
                + self.d4.trainable_variables + [self.gamma_kernel, self.gamma_bias, self.beta_kernel, self.beta_bias])

    def build(self, B, H, W):
        shapes = [(B, H, W, 6), (B, H // 2, W // 2, 64), (B, H // 4, W // 4, 128), (B, H // 8, W // 8, 256)]
        for d, shp in zip((self.d1, self.d2, self.d3, self.d4), shapes):
            if d.kernel is None:
                d.build(shp)
        flat = (H // 8) * (W // 8) * 512
        if self.gamma_kernel is None:
            lim = math.sqrt(6.0 / (flat + 1))
            self.gamma_kernel = torch.empty((flat, 1), dtype=torch.float32).uniform_(-lim, lim).to(self.device)
            self.beta_kernel = torch.empty((flat, 1), dtype=torch.float32).uniform_(-lim, lim).to(self.device)
            self.gamma_bias = torch.zeros(1, dtype=torch.float32, device=self.device)
            self.beta_bias = torch.zeros(1, dtype=torch.float32, device=self.device)

    def _heads(self):
        vs = (self.gamma_kernel, self.gamma_bias, self.beta_kernel, self.beta_bias)
        key = tuple((v.data_ptr(), v._version) for v in vs)
        if key != self._gb_key:
            self._gb_kernel = torch.cat([self.gamma_kernel, self.beta_kernel], dim=1).contiguous()     # [flat, 2]
            self._gb_bias = torch.cat([self.gamma_bias, self.beta_bias]).contiguous()
            self._gb_key = key
        return self._gb_kernel, self._gb_bias

    def call(self, x, actv_map, training="training", *, x_max=None, log_compress_tiled=False, return_heads=False):
        """x: sun-position map [B,H,W,1] (already divided by its maximum, or raw with the device scalar `x_max`);
        actv_map: [B,H,W,8] (6 channels + 2 of padding, from sky_sunrad_input).  Returns (sun_rad [B,H,W,1], gamma_in, beta_in);
        with log_compress_tiled=True the first item is hdr_logCompression(tile(sun_rad, 3)) [B,H,W,3] instead (what
        inference.py:104-105 builds from it)."""
        x = _require_cuda(x, "x")
        B, H, W = x.shape[:3]
        self.build(B, H, W)
        d1 = self.d1(actv_map, training)                     # sunrad_net.py:48-51
        d2 = self.d2(d1, training)
        d3 = self.d3(d2, training)
        d4 = self.d4(d3, training)
        flat = d4.reshape(B, -1)                              # :53 Keras Flatten, (h, w, c) order
        Wgb, bgb = self._heads()
        gb = torch.empty((B, 2), dtype=torch.float32, device=x.device)
        check(LIB.sky_dense_fwd(flat.data_ptr(), Wgb.data_ptr(), bgb.data_ptr(), gb.data_ptr(), B, flat.shape[1], 2, 0, _stream()))   # :54-55
        if x_max is None:
            x_max = torch.ones(1, dtype=torch.float32, device=x.device)
        out3 = torch.empty((B, H, W, 3), dtype=torch.float32, device=x.device)
        lin = None if log_compress_tiled else torch.empty((B, H, W, 1), dtype=torch.float32, device=x.device)
        check(LIB.sky_sun_radiance(x.data_ptr(), x_max.data_ptr(), gb.data_ptr(), out3.data_ptr(), _ptr(lin), B, H * W, self.epsilon,
                                   _stream()))               # :57-69
        gamma_in = beta_in = None
        if return_heads:      # for inspection only (:57-60; inference.py:103 drops them); the kernel applies the sigmoids itself
            gamma_in = torch.sigmoid(gb[:, 0]).view(-1, 1, 1, 1)
            beta_in = torch.sigmoid(gb[:, 1]).view(-1, 1, 1, 1)
        return (out3 if log_compress_tiled else lin), gamma_in, beta_in

    __call__ = call
