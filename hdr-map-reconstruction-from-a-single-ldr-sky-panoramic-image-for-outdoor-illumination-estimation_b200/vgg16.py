"""Host-side mirror of the reference's vgg16.py: ``Vgg16(vgg16_npy_path=None, VGG_MEAN)`` with ``call(bgr) -> (pool1, pool2, pool3)``
(vgg16.py:88-165): scale by 255, subtract the channel means, conv1_1 .. conv3_3 (3x3 SAME + bias + ReLU in the conv epilogue) and three
2x2 max-pools.  `data_dict` maps layer names to (kernel [3,3,C,F], bias) exactly like the npy file the reference loads (vgg16.py:99);
without a file (vgg16.npy is not part of the repository) He-normal weights are drawn so the graph can be timed."""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from ._lib import LIB, check
from .distortion_aware_ops import _require_cuda, _stream
from .sunpose_net import maxpool2d

LAYERS = (("conv1_1", 3, 64), ("conv1_2", 64, 64), ("conv2_1", 64, 128), ("conv2_2", 128, 128), ("conv3_1", 128, 256),
          ("conv3_2", 256, 256), ("conv3_3", 256, 256))


def random_data_dict(seed=0):
    rng = np.random.default_rng(seed)
    return {name: ((rng.standard_normal((3, 3, c, f)) * np.sqrt(2.0 / (9 * c))).astype(np.float32), np.zeros(f, np.float32))
            for name, c, f in LAYERS}


class Vgg16:
    def __init__(self, vgg16_npy_path=None, VGG_MEAN=[103.939, 116.779, 123.68], *, data_dict=None, math_mode=None, device="cuda"):
        if data_dict is None:
            data_dict = (np.load(vgg16_npy_path, encoding='latin1', allow_pickle=True).item() if vgg16_npy_path else random_data_dict())
        self.VGG_MEAN = VGG_MEAN
        self.device = torch.device(device)
        for name, c, f in LAYERS:
            w, b = data_dict[name]
            w = np.asarray(w, np.float32)
            if c == 3:      # the preprocessing kernel writes 4 channels (the 4th zero): pad the kernel's input-channel axis to match
                w = np.concatenate([w, np.zeros((3, 3, 1, f), np.float32)], axis=2)
            layer = ops.conv2d(output_channels=f, k_h=3, k_w=3, strides=1, kernel_initializer=w, bias_initializer=np.asarray(b, np.float32),
                               math_mode=math_mode, device=device)
            layer.build((1, 1, 1, w.shape[2]))
            setattr(self, name, layer)

    def call(self, bgr, training="training"):
        x = _require_cuda(bgr, "bgr")
        B, H, W, _ = x.shape
        pre = torch.empty((B, H, W, 4), dtype=torch.float32, device=x.device)
        check(LIB.sky_vgg_preprocess(x.data_ptr(), pre.data_ptr(), B * H * W, *[float(m) for m in self.VGG_MEAN], _stream()))   # :136-144
        relu = dict(relu=True)
        pool1 = maxpool2d(self.conv1_2(self.conv1_1(pre, **relu), **relu))                       # :146-148
        pool2 = maxpool2d(self.conv2_2(self.conv2_1(pool1, **relu), **relu))                     # :150-152
        pool3 = maxpool2d(self.conv3_3(self.conv3_2(self.conv3_1(pool2, **relu), **relu), **relu))   # :154-157
        return pool1, pool2, pool3

    __call__ = call

    # ---- perceptual-loss direction (train.py:306-312): features of [prediction; target] in one batch, data gradient for the first half ----
    def forward_saved(self, bgr):
        x = _require_cuda(bgr, "bgr")
        B, H, W, _ = x.shape
        pre = torch.empty((B, H, W, 4), dtype=torch.float32, device=x.device)
        check(LIB.sky_vgg_preprocess(x.data_ptr(), pre.data_ptr(), B * H * W, *[float(m) for m in self.VGG_MEAN], _stream()))
        relu = dict(relu=True)
        c11 = self.conv1_1(pre, **relu); c12 = self.conv1_2(c11, **relu); p1 = maxpool2d(c12)
        c21 = self.conv2_1(p1, **relu); c22 = self.conv2_2(c21, **relu); p2 = maxpool2d(c22)
        c31 = self.conv3_1(p2, **relu); c32 = self.conv3_2(c31, **relu); c33 = self.conv3_3(c32, **relu); p3 = maxpool2d(c33)
        self._saved = (pre, c11, c12, p1, c21, c22, p2, c31, c32, c33, p3)
        return p1, p2, p3

    def perceptual_backward(self, n_pred, scale, acc3):
        """The saved batch holds n_pred predictions followed by n_pred targets.  acc3 (3 zeroed fp64) receives sum |pool_i(pred) -
        pool_i(target)|; returns d(scale * sum_i mean |.|) / d(preprocessed prediction) [n_pred,H,W,4] (vgg16.py is frozen: data
        gradients only, each conv's ReLU mask in the epilogue of the data gradient above it)."""
        pre, c11, c12, p1, c21, c22, p2, c31, c32, c33, p3 = self._saved
        n = n_pred
        st = _stream()

        def l1(pool, g, i, accumulate):
            a, b = pool[:n], pool[n:2 * n]
            check(LIB.sky_l1_bwd(a.data_ptr(), b.data_ptr(), g.data_ptr(), acc3[i:i + 1].data_ptr(), a.numel(), float(scale) / a.numel(),
                                 int(accumulate), st))

        def pool_bwd(x, dy):
            B, h, w, C = x.shape
            dx = torch.empty_like(x)
            check(LIB.sky_maxpool2x2_bwd_relu(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), B, h, w, C, 1, st))
            return dx

        g = torch.empty_like(p3[:n])
        l1(p3, g, 2, False)
        g = pool_bwd(c33[:n], g)
        g = self.conv3_3.backward_data(c32[:n], g, mask_src=c32[:n], slope=0.0)
        g = self.conv3_2.backward_data(c31[:n], g, mask_src=c31[:n], slope=0.0)
        g = self.conv3_1.backward_data(p2[:n], g)
        l1(p2, g, 1, True)
        g = pool_bwd(c22[:n], g)
        g = self.conv2_2.backward_data(c21[:n], g, mask_src=c21[:n], slope=0.0)
        g = self.conv2_1.backward_data(p1[:n], g)
        l1(p1, g, 0, True)
        g = pool_bwd(c12[:n], g)
        g = self.conv1_2.backward_data(c11[:n], g, mask_src=c11[:n], slope=0.0)
        return self.conv1_1.backward_data(pre[:n], g)
