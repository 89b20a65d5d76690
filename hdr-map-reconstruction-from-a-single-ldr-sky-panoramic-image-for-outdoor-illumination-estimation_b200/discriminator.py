"""Host-side mirror of the reference's discriminator.py: ``model(im_height, im_width)`` with ``call([ldr, hdr], training)``
(discriminator.py:29-50): concat -> four `downsampling` blocks (the same class as sunrad_net.py:8-28) -> Conv2D(1, 4, VALID) + bias.
Inference mode (training=False, as the generator step calls it, train.py:300): BatchNormalization folded into the convs.  The VALID
4x4 output conv is the SAME conv cropped: VALID[i, j] == SAME[i + 1, j + 1] (SAME pads one row / column in front for k = 4).
Training mode (discriminator_in_step, train.py:351-380): the real and the generated pair run as ONE batch of 2B whose halves are
normalised separately (two Keras calls = two sets of batch statistics), which halves the launches and lets one weight-gradient launch
per layer cover both calls.  ``infer_forward`` / ``infer_backward_data`` give the generator its adversarial gradient (LSGAN gen_loss,
train.py:327) through the folded convs."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import LIB, check
from .distortion_aware_ops import _MATH, DEFAULT_MATH_MODE, _require_cuda, _stream
from .ops import TransposedPack, conv_backward_data, conv_backward_filter
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
        self._tp_out = None
        self._saved = None

    def _invalidate(self):
        self._key = None
        if self._tp_out is not None:
            self._tp_out.key = None

    def owner_list(self):
        out = []
        for d in (self.d1, self.d2, self.d3, self.d4):
            out += [(d, "kernel")] + ([(d, "gamma"), (d, "beta")] if d.apply_norm else [])
        return out + [(self, "out_kernel"), (self, "out_bias")]

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

    @staticmethod
    def concat(ldr, hdr, out=None):
        """tf.concat([ldr, hdr], -1) (discriminator.py:42), zero-padded to 8 channels so the first conv gathers 16-byte chunks."""
        B, H, W, _ = ldr.shape
        cat = out if out is not None else torch.empty((B, H, W, 8), dtype=torch.float32, device=ldr.device)
        check(LIB.sky_concat2_pad(ldr.data_ptr(), 3, hdr.data_ptr(), 3, cat.data_ptr(), 8, B * H * W, _stream()))
        return cat

    def _out_conv(self, h):
        """Conv2D(1, 4, VALID) + bias (discriminator.py:48) as the SAME map [B,hh,ww,1]; its window [1:hh-2, 1:ww-2] is the VALID output."""
        key = (self.out_kernel.data_ptr(), self.out_kernel._version)
        mode = _MATH[self.math_mode]
        if key != self._key:
            if self._packed is None:
                self._packed = torch.empty(LIB.sky_da_packed_weight_bytes(512, 1, 4, mode), dtype=torch.uint8, device=self.device)
            check(LIB.sky_da_pack_weights(self.out_kernel.data_ptr(), self._packed.data_ptr(), 512, 1, 4, mode, _stream()))
            self._key = key
        B, hh, ww, _ = h.shape
        same = torch.empty((B, hh, ww, 1), dtype=torch.float32, device=h.device)
        check(LIB.sky_conv2d_fwd(h.data_ptr(), self._packed.data_ptr(), self.out_bias.data_ptr(), same.data_ptr(), None, None, B, hh, ww,
                                 512, 1, 4, 1, _lib.EPI_NONE, 0.0, mode, _stream()))
        return same

    def call(self, x, training="training"):
        ldr, hdr = (_require_cuda(t, "x") for t in x)
        B, H, W, _ = ldr.shape
        self.build(B, H, W)
        cat = self.concat(ldr, hdr)                                                                                    # :42
        h = self.d4(self.d3(self.d2(self.d1(cat, training), training), training), training)                            # :43-46
        same = self._out_conv(h)                                                                                       # :48
        hh, ww = same.shape[1], same.shape[2]
        return same[:, 1:hh - 2, 1:ww - 2, :]          # VALID window: rows 0..h-4 of the VALID output == rows 1..h-3 of the SAME output

    __call__ = call

    # ---- inference-mode pass with its data gradient (generator step, train.py:300, 327) -------------------------------------------
    def infer_forward(self, cat8):
        """cat8: concat(ldr, y_final_lin) [B,H,W,8].  Returns the SAME map of the output conv; activations are kept."""
        B, H, W, _ = cat8.shape
        self.build(B, H, W)
        a1 = self.d1(cat8, False)
        a2 = self.d2(a1, False)
        a3 = self.d3(a2, False)
        a4 = self.d4(a3, False)
        same = self._out_conv(a4)
        self._saved_infer = (cat8, a1, a2, a3, a4)
        return same

    def _out_backward_data(self, g_same, a4):
        if self._tp_out is None:
            self._tp_out = TransposedPack(4, 512, 1, self.math_mode, self.device)
        return conv_backward_data(self._tp_out, self.out_kernel, tuple(a4.shape), g_same, 1, mask_src=a4, slope=self.d4.alpha)

    def infer_backward_data(self, g_same):
        """g_same: gradient w.r.t. the SAME map (zero outside the VALID window).  Returns the gradient w.r.t. cat8 [B,H,W,8]."""
        cat8, a1, a2, a3, a4 = self._saved_infer
        g = self._out_backward_data(g_same, a4)                                  # masked by lrelu'(a4)
        g = self.d4.infer_backward_data(a3.shape, g, dx_mask_src=a3)
        g = self.d3.infer_backward_data(a2.shape, g, dx_mask_src=a2)
        g = self.d2.infer_backward_data(a1.shape, g, dx_mask_src=a1)
        return self.d1.infer_backward_data(cat8.shape, g)

    # ---- training-mode pass over the (real, generated) pair (discriminator_in_step, train.py:351-380) ----------------------------
    def train_forward(self, cat16, groups=2):
        """cat16: [2B,H,W,8], first half concat(ldr, hdr_t), second half concat(ldr, y_final_lin).  Returns the SAME map [2B,hh,ww,1]."""
        B2, H, W, _ = cat16.shape
        self.build(B2, H, W)
        a1 = self.d1.train_forward(cat16, groups)
        a2 = self.d2.train_forward(a1, groups)
        a3 = self.d3.train_forward(a2, groups)
        a4 = self.d4.train_forward(a3, groups)
        self._saved = a4
        return self._out_conv(a4)

    def train_backward(self, g_same, fv):
        """g_same [2B,hh,ww,1]: gradient of total_disc_loss w.r.t. the SAME map.  Variable gradients go to fv."""
        a4 = self._saved
        acc = bool(getattr(fv, "zeroed_each_step", False))       # the caller zeroed fv.flat_g: the conv weight gradients add to it
        conv_backward_filter(a4, g_same, 4, 1, fv.grad(self, "out_kernel").view(-1, 1), fv.grad(self, "out_bias"), acc)
        # the BN backward applies the LeakyReLU mask itself: the out conv's data gradient must arrive unmasked
        if self._tp_out is None:
            self._tp_out = TransposedPack(4, 512, 1, self.math_mode, self.device)
        g = conv_backward_data(self._tp_out, self.out_kernel, tuple(a4.shape), g_same, 1)
        grads = lambda d: {k: fv.grad(d, k) for k in (("kernel", "gamma", "beta") if d.apply_norm else ("kernel",))}
        g = self.d4.train_backward(g, grads(self.d4), accumulate=acc)
        g = self.d3.train_backward(g, grads(self.d3), accumulate=acc)
        g = self.d2.train_backward(g, grads(self.d2), dx_mask_src=self.d1._saved["a"], accumulate=acc)
        self.d1.train_backward(g, grads(self.d1), need_dx=False, accumulate=acc)
