"""Host-side mirror of the reference's sunpose_net.py: ``sunposeLayer(filter_out, k_h, k_w, strides, dilation_rate)`` and
``model(im_height, im_width, da_kernel_size, dilation_rate)`` with ``sunposeEstimation(x) -> (softmax, [act1, act2, act3])``.

`distortion_aware=True` (default) wires the convolutions the way the commented lines sunpose_net.py:11,16 do —
``distortion_aware_ops.conv2d(filter_out, kernel_size=k_h, strides=strides, dilation_rate=dilation_rate)`` — which is the
north-star path; `distortion_aware=False` gives the plain ``ops.conv2d`` wiring that is live in the committed reference."""
from __future__ import annotations

import math
import os

import torch

from . import _lib, ops
from ._lib import LIB, check
from .distortion_aware_ops import _initializer, _require_cuda, _stream, conv2d as da_conv2d, conv2d_backward, zero_
from .generator import InstanceNormalization


def maxpool2d(x):
    """ops.maxpool2d(kernel_size=2).call (ops.py:299-300): tf.nn.max_pool 2x2 / 2 SAME."""
    x = _require_cuda(x, "x")
    B, h, w, C = x.shape
    y = torch.empty((B, (h + 1) // 2, (w + 1) // 2, C), dtype=torch.float32, device=x.device)
    check(LIB.sky_maxpool2x2_fwd(x.data_ptr(), y.data_ptr(), B, h, w, C, _stream()))
    return y


DENSE_BWD_KERNEL = os.environ.get("SKY_DENSE_BWD_KERNEL", "nt")      # "t": transposed copy + the forward kernel (round 1)


class Dense:
    """tf.keras.layers.Dense(units): kernel [in, units] glorot_uniform, bias zeros."""

    def __init__(self, units, device="cuda"):
        self.units = int(units)
        self.device = torch.device(device)
        self.kernel = None
        self.bias = None
        self._kernel_t = None
        self._kernel_t_key = None

    def build(self, input_shape):
        k = int(input_shape[-1])
        self.kernel = _initializer("glorot_uniform", (k, self.units), k, self.units, self.device)
        self.bias = torch.zeros(self.units, dtype=torch.float32, device=self.device)

    def _invalidate(self):
        self._kernel_t_key = None

    def kernel_transposed(self):
        """W^T [units, in], rebuilt only when the variable changes: the backward streams it like the forward streams W."""
        key = (self.kernel.data_ptr(), self.kernel._version)
        if self._kernel_t_key != key:
            k, n = self.kernel.shape
            if self._kernel_t is None or tuple(self._kernel_t.shape) != (n, k):
                self._kernel_t = torch.empty((n, k), dtype=torch.float32, device=self.device)
            check(LIB.sky_transpose(self.kernel.data_ptr(), self._kernel_t.data_ptr(), k, n, _stream()))
            self._kernel_t_key = key
        return self._kernel_t

    def backward_data(self, dy, act=None):
        """dx = dy . W^T, zeroed where `act` (the ReLU output that fed this layer) is not positive."""
        dy = _require_cuda(dy, "dy")
        B, n = dy.shape
        k = self.kernel.shape[0]
        dx = torch.empty((B, k), dtype=torch.float32, device=dy.device)
        if DENSE_BWD_KERNEL == "nt" and n % 4 == 0 and k >= 256:
            # W streamed in its own [in, units] layout: no transposed copy per step
            check(LIB.sky_dense_bwd_data_nt(dy.data_ptr(), self.kernel.data_ptr(), None if act is None else act.data_ptr(), dx.data_ptr(),
                                            B, k, n, _stream()))
        else:
            check(LIB.sky_dense_bwd_data(dy.data_ptr(), self.kernel_transposed().data_ptr(), None if act is None else act.data_ptr(),
                                         dx.data_ptr(), B, k, n, _stream()))
        return dx

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


def maxpool2d_backward(x, dy):
    """Gradient of ops.maxpool2d w.r.t. its input `x` (TensorFlow MaxPoolGrad: first maximum of each window)."""
    x, dy = _require_cuda(x, "x"), _require_cuda(dy, "dy")
    B, h, w, C = x.shape
    dx = torch.empty_like(x)
    check(LIB.sky_maxpool2x2_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), B, h, w, C, _stream()))
    return dx


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
            self._stats = torch.empty(2, B, self.filter_out, 2, dtype=torch.float64, device=x.device)
        zero_(self._stats)
        conv1 = self.conv1(x, stats=self._stats[0])                                  # :21
        actv1 = self.norm1.apply(conv1, self._stats[0], leaky_slope=0.0)             # :22-23  relu == leaky_relu(0)
        conv2 = self.conv2(actv1, stats=self._stats[1])                              # :25
        actv2 = self.norm2.apply(conv2, self._stats[1], leaky_slope=0.0)             # :26-29
        self._saved = (x, conv1, actv1, conv2, actv2)
        return actv2

    __call__ = call

    def _conv_backward_data(self, conv, x, dy):
        if isinstance(conv, da_conv2d):
            return conv2d_backward(conv, x, dy, need_dw=False)[0]
        return conv.backward_data(x, dy)

    def backward_data(self, g_out, need_input_grad=True):
        """Gradient w.r.t. the layer input given the gradient w.r.t. its output (the Grad-CAM activation map): ReLU and
        instance-norm backward fused (sky_instnorm_bwd, slope 0), then the conv data gradient; weights are constants."""
        x, conv1, actv1, conv2, actv2 = self._saved
        B, h, w, F = conv2.shape
        if getattr(self, "_bwd_scratch", None) is None or self._bwd_scratch[0].shape[0] != B:
            self._bwd_scratch = (torch.empty(B, F, 2, dtype=torch.float64, device=x.device),
                                 torch.zeros(2, F, dtype=torch.float32, device=x.device))
        sums, dgb = self._bwd_scratch

        def norm_bwd(norm, pre, stats, dy, act):
            dx = torch.empty_like(pre)
            check(LIB.sky_instnorm_bwd(pre.data_ptr(), stats.data_ptr(), norm.gamma.data_ptr(), dy.data_ptr(), act.data_ptr(), None,
                                       sums.data_ptr(), dx.data_ptr(), dgb[0].data_ptr(), dgb[1].data_ptr(), B, h, w, F,
                                       norm.epsilon, 0.0, _stream()))
            return dx

        g = norm_bwd(self.norm2, conv2, self._stats[1], _require_cuda(g_out, "g_out"), actv2)
        g = self._conv_backward_data(self.conv2, actv1, g)
        if not need_input_grad:
            return None
        g = norm_bwd(self.norm1, conv1, self._stats[0], g, actv1)
        return self._conv_backward_data(self.conv1, x, g)


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
        self._saved = None

    @property
    def trainable_variables(self):
        return (self.sunlayer1.trainable_variables + self.sunlayer2.trainable_variables + self.sunlayer3.trainable_variables
                + self.fc1.trainable_variables + self.fc2.trainable_variables)

    def set_weights(self, w):
        """w: sunlayer{1,2,3} -> dict conv{1,2}_kernel [k*k*C, F] / conv{1,2}_bias / norm{1,2}_gamma / norm{1,2}_beta, fc{1,2} ->
        (kernel [in, units], bias); layers must have been built (one sunposeEstimation call)."""
        for name in ("sunlayer1", "sunlayer2", "sunlayer3"):
            layer, d = getattr(self, name), w[name]
            for i, (conv, norm) in enumerate(((layer.conv1, layer.norm1), (layer.conv2, layer.norm2)), start=1):
                kern, bias = torch.as_tensor(d[f"conv{i}_kernel"]), torch.as_tensor(d[f"conv{i}_bias"])
                if isinstance(conv, da_conv2d):
                    conv.kernel.copy_(kern)
                    conv.bias.copy_(bias)
                else:
                    conv.w.copy_(kern.reshape(conv.w.shape))
                    conv.biases.copy_(bias)
                norm.gamma.copy_(torch.as_tensor(d[f"norm{i}_gamma"]))
                norm.beta.copy_(torch.as_tensor(d[f"norm{i}_beta"]))
        for name in ("fc1", "fc2"):
            getattr(self, name).kernel.copy_(torch.as_tensor(w[name][0]))
            getattr(self, name).bias.copy_(torch.as_tensor(w[name][1]))

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
        self._saved = (sm, actv1_s, actv2_s, tuple(pool3_s.shape), [sunlayer1, sunlayer2, sunlayer3])
        self._saved_flat = flat
        return sm, [sunlayer1, sunlayer2, sunlayer3]

    def class_score(self, sm, sunpose_gt=None):
        """y_c = tf.math.reduce_max(sunpose_cmf, axis=1) (inference.py:98) — or, with `sunpose_gt`, the softmax entry at
        argmax(sunpose_gt) (train.py:263-265) — as a handle that can be differentiated with respect to the three activation maps:
        what grad_cam.layer's tf.gradients(y_c, A_k) (grad_cam.py:31) asks for."""
        if self._saved is None or sm is not self._saved[0]:
            raise ValueError("class_score() takes the softmax returned by the last sunposeEstimation() call")
        return ClassScore(self, sunpose_gt)


class ClassScore:
    """max_j softmax_j per panorama plus the backward sweep softmax -> fc2 -> fc1 -> pool3 -> sunlayer3 -> pool2 -> sunlayer2
    -> pool1 that yields d(sum_b y_c[b]) / dA_k for the three maps (run once, on first use).  TensorFlow would run one
    backward sub-graph per tf.gradients call; the three share every node, so one sweep gives identical values."""

    def __init__(self, net, sunpose_gt=None):
        self._net = net
        sm, actv1_s, actv2_s, pool_shape, acts = net._saved
        B, n = sm.shape
        self.value = torch.empty(B, dtype=torch.float32, device=sm.device)
        self._seed = torch.empty_like(sm)
        if sunpose_gt is None:
            check(LIB.sky_softmax_max_bwd(sm.data_ptr(), actv2_s.data_ptr(), self.value.data_ptr(), self._seed.data_ptr(), B, n, _stream()))
        else:
            gt = _require_cuda(sunpose_gt, "sunpose_gt")
            self.pick = torch.empty(B, dtype=torch.int32, device=sm.device)
            check(LIB.sky_argmax_rows(gt.data_ptr(), self.pick.data_ptr(), B, n, _stream()))
            check(LIB.sky_softmax_pick_bwd(sm.data_ptr(), actv2_s.data_ptr(), self.pick.data_ptr(), self.value.data_ptr(),
                                           self._seed.data_ptr(), B, n, _stream()))
        self._grads = None

    def _sweep(self):
        net = self._net
        sm, actv1_s, actv2_s, pool_shape, acts = net._saved
        g = net.fc2.backward_data(self._seed, act=actv1_s)          # through fc2, then the ReLU of sunpose_net.py:66
        g = net.fc1.backward_data(g)                                 # through fc1 -> Flatten
        g3 = maxpool2d_backward(acts[2], g.view(pool_shape))         # pool3_s
        g = net.sunlayer3.backward_data(g3)
        g2 = maxpool2d_backward(acts[1], g)                          # pool2_s
        g = net.sunlayer2.backward_data(g2)
        g1 = maxpool2d_backward(acts[0], g)                          # pool1_s
        self._grads = [g1, g2, g3]

    def gradient(self, A_k):
        if self._grads is None:
            self._sweep()
        for a, g in zip(self._net._saved[4], self._grads):
            if a is A_k:
                return g
        raise ValueError("A_k must be one of the activation maps returned by sunposeEstimation()")
