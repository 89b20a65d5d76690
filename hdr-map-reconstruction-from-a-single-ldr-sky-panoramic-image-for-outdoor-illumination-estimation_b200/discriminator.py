"""Host-side mirror of the reference's discriminator.py: ``model(im_height, im_width)`` with ``call([ldr, hdr], training)``
(discriminator.py:29-50): concat -> four `downsampling` blocks (the same class as sunrad_net.py:8-28) -> Conv2D(1, 4, VALID) + bias.
Inference mode (training=False, as the generator step calls it, train.py:300): BatchNormalization folded into the convs.  The VALID
4x4 output conv is the SAME conv cropped: VALID[i, j] == SAME[i + 1, j + 1] (SAME pads one row / column in front for k = 4)."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import LIB, check
from .distortion_aware_ops import _MATH, DEFAULT_MATH_MODE, _require_cuda, _stream
from .sunrad_net import downsampling


class model:
    def __init__(self, im_height=32, im_width=128, da_kernel_size=3, dilation_rate=1, *, math_mode=None, device="cuda"):
        kw = dict(math_mode=math_mode, device=device)
        self.math_mode = math_mode or DEFAULT_MATH_MODE
        self.device = torch.device(device)
        self.d1 = downsampling(64, 4, strides=2, apply_norm=False, channel_pad=8, **kw)     # discriminator.py:33
        self.d2 = downsampling(128, 4, strides=2, apply_norm=True, **kw)
        self.d3 = downsampling(256, 4, strides=2, apply_norm=True, **kw)
        self.d4 = downsampling(512, 4, strides=1, apply_norm=True, **kw)
        self.out_kernel = self.out_bias = None                                              # Conv2D(1, 4), N(0, 0.02), bias zeros (:38-39)
        self._packed = None
        self._key = None

    def build(self, B, H, W):
        shapes = [(B, H, W, 6), (B, H // 2, W // 2, 64), (B, H // 4, W // 4, 128), (B, H // 8, W // 8, 256)]
        for d, shp in zip((self.d1, self.d2, self.d3, self.d4), shapes):
            if d.kernel is None:
                d.build(shp)
        if self.out_kernel is None:
            self.out_kernel = torch.empty((4, 4, 512, 1), dtype=torch.float32).normal_(0.0, 0.02).to(self.device)
            self.out_bias = torch.zeros(1, dtype=torch.float32, device=self.device)

    @property
    def trainable_variables(self):
        return (self.d1.trainable_variables + self.d2.trainable_variables + self.d3.trainable_variables + self.d4.trainable_variables
                + [self.out_kernel, self.out_bias])

    def set_weights(self, w):
        """w: d1..d4 -> dict(kernel (, gamma, beta, moving_mean, moving_variance)), out -> (kernel [4,4,512,1], bias [1])."""
        for name in ("d1", "d2", "d3", "d4"):
            for key, val in w[name].items():
                getattr(getattr(self, name), key).copy_(torch.as_tensor(val))
        self.out_kernel.copy_(torch.as_tensor(w["out"][0]))
        self.out_bias.copy_(torch.as_tensor(w["out"][1]))

    def call(self, x, training="training"):
        ldr, hdr = (_require_cuda(t, "x") for t in x)
        B, H, W, _ = ldr.shape
        self.build(B, H, W)
        cat = torch.empty((B, H, W, 8), dtype=torch.float32, device=ldr.device)
        check(LIB.sky_concat2_pad(ldr.data_ptr(), 3, hdr.data_ptr(), 3, cat.data_ptr(), 8, B * H * W, _stream()))      # :42
        h = self.d4(self.d3(self.d2(self.d1(cat, training), training), training), training)                            # :43-46
        key = (self.out_kernel.data_ptr(), self.out_kernel._version)
        mode = _MATH[self.math_mode]
        if key != self._key:
            self._packed = torch.empty(LIB.sky_da_packed_weight_bytes(512, 1, 4, mode), dtype=torch.uint8, device=self.device)
            check(LIB.sky_da_pack_weights(self.out_kernel.data_ptr(), self._packed.data_ptr(), 512, 1, 4, mode, _stream()))
            self._key = key
        _, hh, ww, _ = h.shape
        same = torch.empty((B, hh, ww, 1), dtype=torch.float32, device=ldr.device)
        check(LIB.sky_conv2d_fwd(h.data_ptr(), self._packed.data_ptr(), self.out_bias.data_ptr(), same.data_ptr(), None, None, B, hh, ww,
                                 512, 1, 4, 1, _lib.EPI_NONE, 0.0, mode, _stream()))                                   # :48
        return same[:, 1:hh - 2, 1:ww - 2, :]          # VALID window: rows 0..h-4 of the VALID output == rows 1..h-3 of the SAME output

    __call__ = call
