"""Host-side mirror of the reference's sunrad_net.py: ``downsampling(filters, kernel_size, strides=2, apply_norm=True)`` and
``sunRadNet(epsilon=1e-5)`` with ``call(x, actv_map, training) -> (sun_rad, gamma_in, beta_in)`` (sunrad_net.py:8-71).

Inference mode (training=False, inference.py:103): Keras BatchNormalization uses its moving statistics, so each
conv -> BN -> LeakyReLU(0.3) block is ONE conv launch: the BN scale is folded into the (bias-free) kernel and its shift into a
bias by ``sky_bn_fold`` whenever a variable changes, and LeakyReLU rides in the conv epilogue.  d4 has 512 filters and runs as
two 256-filter slices inside ``sky_conv2d_fwd``.  The two Dense(1) heads are one [32768, 2] weight-streaming product, and the
Dirac-delta-like radiance function, its 30000 clamp, hdr_logCompression and the x3 tile are one elementwise kernel.
Training mode (training=True, train.py:288 via generator.sun_rad_estimation): conv (unfolded kernel) -> batch statistics
(``sky_bn_train_stats``, which also steps the moving statistics) -> normalise + LeakyReLU (``sky_bn_train_apply``), every
intermediate kept for ``train_backward``: BN backward fused with the LeakyReLU mask (``sky_bn_train_bwd``), weight gradient
(``sky_conv2d_bwd_filter``) and data gradient (``sky_conv2d_bwd_data``, a gather over dy with the transposed kernel).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib
from ._lib import LIB, check
from .distortion_aware_ops import _MATH, DEFAULT_MATH_MODE, _ptr, _require_cuda, _stream
from .ops import TransposedPack, conv_backward_data, conv_backward_filter


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
        self.momentum = 0.99                    # Keras BatchNormalization default
        self._padded = self._packed_train = None
        self._train_key = None
        self._tp_train = self._tp_fold = None   # transposed packs of the padded raw / BN-folded kernel (data gradients)
        self._saved = None

    def _invalidate(self):
        self._key = self._train_key = None
        for tp in (self._tp_train, self._tp_fold):
            if tp is not None:
                tp.key = None

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

    # ---- training mode ----------------------------------------------------------------------------------------------------------
    def _cp(self):
        return max(self._channels, self.channel_pad or self._channels)

    def _train_weights(self):
        """(padded raw kernel [k,k,Cp,F], its packed image) — no BatchNormalization fold in training mode."""
        key = (self.kernel.data_ptr(), self.kernel._version)
        if key != self._train_key:
            k, C, F, Cp = self.kernel_size, self._channels, self.filters, self._cp()
            if self._padded is None:
                self._padded = torch.zeros((k, k, Cp, F), dtype=torch.float32, device=self.device) if Cp != C else None
                nbytes = LIB.sky_da_packed_weight_bytes(Cp, F, k, _MATH[self.math_mode])
                self._packed_train = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            src = self.kernel
            if Cp != C:
                self._padded[:, :, :C, :].copy_(self.kernel)
                src = self._padded
            check(LIB.sky_da_pack_weights(src.data_ptr(), self._packed_train.data_ptr(), Cp, F, k, _MATH[self.math_mode], _stream()))
            self._train_key = key
        return (self.kernel if self._padded is None else self._padded), self._packed_train

    def train_forward(self, x, groups=1, update_moving=True):
        """Conv -> BatchNormalization(batch statistics, per batch group) -> LeakyReLU with every intermediate saved."""
        x = _require_cuda(x, "x")
        if self.kernel is None:
            raise RuntimeError("downsampling.train_forward: build() the layer first (the owning model does)")
        B, h, w, Cx = x.shape
        if Cx != self._cp():
            raise ValueError(f"input has {Cx} channels, the layer expects {self._cp()} (channel-padded)")
        s, F, k = self.strides, self.filters, self.kernel_size
        oh, ow = -(-h // s), -(-w // s)
        _, packed = self._train_weights()
        mode = _MATH[self.math_mode]
        if not self.apply_norm:
            a = torch.empty((B, oh, ow, F), dtype=torch.float32, device=x.device)
            check(LIB.sky_conv2d_fwd(x.data_ptr(), packed.data_ptr(), None, a.data_ptr(), None, None, B, h, w, Cx, F, k, s,
                                     _lib.EPI_LEAKY_RELU, self.alpha, mode, _stream()))
            self._saved = dict(x=x, z=None, a=a, groups=groups)
            return a
        z = torch.empty((B, oh, ow, F), dtype=torch.float32, device=x.device)
        check(LIB.sky_conv2d_fwd(x.data_ptr(), packed.data_ptr(), None, z.data_ptr(), None, None, B, h, w, Cx, F, k, s, _lib.EPI_NONE, 0.0,
                                 mode, _stream()))
        sums = torch.empty((groups, F, 2), dtype=torch.float64, device=x.device)
        mean_var = torch.empty((groups, F, 2), dtype=torch.float32, device=x.device)
        check(LIB.sky_bn_train_stats(z.data_ptr(), sums.data_ptr(), mean_var.data_ptr(), self.moving_mean.data_ptr() if update_moving else None,
                                     self.moving_variance.data_ptr() if update_moving else None, B, oh * ow, F, groups, self.momentum, _stream()))
        a = torch.empty_like(z)
        check(LIB.sky_bn_train_apply(z.data_ptr(), mean_var.data_ptr(), self.gamma.data_ptr(), self.beta.data_ptr(), a.data_ptr(), B, oh * ow, F,
                                     groups, self.epsilon, _lib.EPI_LEAKY_RELU, self.alpha, _stream()))
        self._saved = dict(x=x, z=z, a=a, mean_var=mean_var, sums=sums, groups=groups)
        self._saved_moving = bool(update_moving)
        return a

    def train_backward(self, g, grads, need_dx=True, dx_mask_src=None, accumulate=False):
        """g: gradient w.r.t. the block output `a` — raw for a BatchNormalization block (the LeakyReLU mask is applied inside the BN
        backward), already multiplied by lrelu'(a) for the norm-free block (the upstream data gradient's SKY_EPI_MASK epilogue does it).
        grads: dict kernel / gamma / beta -> gradient views (gamma / beta accumulate: zero them once per step).  Returns dx (masked by
        dx_mask_src's LeakyReLU gradient when given) or None."""
        sv = self._saved
        x, a = sv["x"], sv["a"]
        B, h, w, Cx = x.shape
        F, k, s = self.filters, self.kernel_size, self.strides
        g = _require_cuda(g, "g")
        if self.apply_norm:
            dz = torch.empty_like(a)
            oh, ow = a.shape[1], a.shape[2]
            check(LIB.sky_bn_train_bwd(sv["z"].data_ptr(), sv["mean_var"].data_ptr(), self.gamma.data_ptr(), g.data_ptr(), a.data_ptr(),
                                       sv["sums"].data_ptr(), dz.data_ptr(), grads["gamma"].data_ptr(), grads["beta"].data_ptr(), B, oh * ow, F,
                                       sv["groups"], self.epsilon, self.alpha, _stream()))
        else:
            dz = g
        conv_backward_filter(x, dz, k, s, grads["kernel"].view(-1, F), None, accumulate, C_store=self._channels)
        if not need_dx:
            return None
        if self._tp_train is None:
            self._tp_train = TransposedPack(k, Cx, F, self.math_mode, self.device)
        wsrc, _ = self._train_weights()
        return conv_backward_data(self._tp_train, wsrc, tuple(x.shape), dz, s, dx_mask_src, self.alpha)

    def infer_backward_data(self, in_shape, g_masked, dx_mask_src=None):
        """Data gradient through the inference-mode block y = lrelu(conv_folded(x) + shift): g_masked = dL/dy * lrelu'(y)."""
        self._packed_weights()                                  # makes sure the folded kernel is current
        Cp = self._cp()
        if self._tp_fold is None:
            self._tp_fold = TransposedPack(self.kernel_size, Cp, self.filters, self.math_mode, self.device)
        return conv_backward_data(self._tp_fold, self._folded, tuple(in_shape), g_masked, self.strides, dx_mask_src, self.alpha)

    def call(self, x, training="training"):
        if training is True or training == "training":
            return self.train_forward(x)
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
        # Dense(1) x 2 (:43-44) held as ONE [flat, 2] kernel / [2] bias (column 0 = gamma head, column 1 = beta head): both heads are
        # one weight-streaming product forward and one kernel backward; gamma_kernel / beta_kernel / *_bias are views
        self.gb_kernel = self.gb_bias = None
        self._saved = None

    @property
    def gamma_kernel(self):
        return None if self.gb_kernel is None else self.gb_kernel[:, 0:1]

    @property
    def beta_kernel(self):
        return None if self.gb_kernel is None else self.gb_kernel[:, 1:2]

    @property
    def gamma_bias(self):
        return None if self.gb_bias is None else self.gb_bias[0:1]

    @property
    def beta_bias(self):
        return None if self.gb_bias is None else self.gb_bias[1:2]

    @property
    def trainable_variables(self):
        return (self.d1.trainable_variables + self.d2.trainable_variables + self.d3.trainable_variables
                + self.d4.trainable_variables + [self.gb_kernel, self.gb_bias])

    def build(self, B, H, W):
        shapes = [(B, H, W, 6), (B, H // 2, W // 2, 64), (B, H // 4, W // 4, 128), (B, H // 8, W // 8, 256)]
        for d, shp in zip((self.d1, self.d2, self.d3, self.d4), shapes):
            if d.kernel is None:
                d.build(shp)
        flat = (H // 8) * (W // 8) * 512
        if self.gb_kernel is None:
            lim = math.sqrt(6.0 / (flat + 1))                 # glorot_uniform of each Dense(1)
            self.gb_kernel = torch.empty((flat, 2), dtype=torch.float32).uniform_(-lim, lim).to(self.device)
            self.gb_bias = torch.zeros(2, dtype=torch.float32, device=self.device)

    def _heads(self):
        return self.gb_kernel, self.gb_bias

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
        self._saved = dict(sm=x, x_max=x_max, flat=flat, d4_shape=tuple(d4.shape))
        Wgb, bgb = self._heads()
        gb = torch.empty((B, 2), dtype=torch.float32, device=x.device)
        check(LIB.sky_dense_fwd(flat.data_ptr(), Wgb.data_ptr(), bgb.data_ptr(), gb.data_ptr(), B, flat.shape[1], 2, 0, _stream()))   # :54-55
        if x_max is None:
            x_max = torch.ones(1, dtype=torch.float32, device=x.device)
        out3 = torch.empty((B, H, W, 3), dtype=torch.float32, device=x.device)
        lin = None if log_compress_tiled else torch.empty((B, H, W, 1), dtype=torch.float32, device=x.device)
        check(LIB.sky_sun_radiance(x.data_ptr(), x_max.data_ptr(), gb.data_ptr(), out3.data_ptr(), _ptr(lin), B, H * W, self.epsilon,
                                   _stream()))               # :57-69
        self._saved.update(x_max=x_max, gb=gb)
        gamma_in = beta_in = None
        if return_heads:      # for inspection only (:57-60; inference.py:103 drops them); the kernel applies the sigmoids itself
            gamma_in = torch.sigmoid(gb[:, 0]).view(-1, 1, 1, 1)
            beta_in = torch.sigmoid(gb[:, 1]).view(-1, 1, 1, 1)
        return (out3 if log_compress_tiled else lin), gamma_in, beta_in

    __call__ = call

    def train_backward(self, d_out3, grads, dsm, accumulate_dsm=True, accumulate_dw=False):
        """Backward of the training-mode call whose first output was hdr_logCompression(tile(sun_rad, 3)): d_out3 [B,H,W,3] is the
        gradient w.r.t. that tensor.  grads: dict d1..d4 -> {kernel (, gamma, beta)}, gb_kernel, gb_bias -> gradient views.  The
        gradient w.r.t. the (un-normalised) sun-position map, max-normalisation of generator.py:160 included, is added to dsm."""
        sv = self._saved
        sm, gb, flat = sv["sm"], sv["gb"], sv["flat"]
        B = sm.shape[0]
        hw = sm.numel() // B
        dnorm = torch.empty((B, hw), dtype=torch.float32, device=sm.device)
        dgb64 = torch.empty((B, 2), dtype=torch.float64, device=sm.device)
        red2 = torch.empty(2, dtype=torch.float64, device=sm.device)
        check(LIB.sky_sun_radiance_bwd(sm.data_ptr(), sv["x_max"].data_ptr(), gb.data_ptr(), d_out3.data_ptr(), dnorm.data_ptr(), dgb64.data_ptr(),
                                       red2.data_ptr(), B, hw, self.epsilon, _stream()))
        check(LIB.sky_maxnorm_bwd(sm.data_ptr(), sv["x_max"].data_ptr(), dnorm.data_ptr(), red2.data_ptr(), dsm.data_ptr(), sm.numel(),
                                  int(bool(accumulate_dsm)), _stream()))
        dflat = torch.empty_like(flat)
        check(LIB.sky_sunrad_heads_bwd(flat.data_ptr(), self.gb_kernel.data_ptr(), dgb64.data_ptr(), grads["gb_kernel"].data_ptr(),
                                       grads["gb_bias"].data_ptr(), dflat.data_ptr(), B, flat.shape[1], _stream()))
        g = dflat.view(sv["d4_shape"])
        # accumulate_dw: the kernel gradient views were zeroed by the caller (the flat buffer, once per step)
        g = self.d4.train_backward(g, grads["d4"], accumulate=accumulate_dw)
        g = self.d3.train_backward(g, grads["d3"], accumulate=accumulate_dw)
        g = self.d2.train_backward(g, grads["d2"], dx_mask_src=self.d1._saved["a"], accumulate=accumulate_dw)      # d1 has no norm: its LeakyReLU mask rides here
        self.d1.train_backward(g, grads["d1"], need_dx=False, accumulate=accumulate_dw)
