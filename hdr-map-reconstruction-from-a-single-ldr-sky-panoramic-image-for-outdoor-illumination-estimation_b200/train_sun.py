"""Host-side mirror of the compute step of the reference's train_sun.py: ``sun_train_step`` (train_sun.py:220-264) — the pre-training
step of the sun-position network (BASELINE configs[1]):

    forward (sunposeEstimation) -> Grad-CAM maps at the ground-truth class (outside the tape, :231-240) -> loss = KLDivergence(gt, sm)
    + sum_l mean |DoG_l(pred) - DoG_l(gt)| (:245-255) -> gradients w.r.t. every variable of the network (:257) -> Adam (:258).

Backward: KL and DoG adjoints -> softmax backward (+ ReLU mask) -> Dense weight / data gradients -> max-pool gradient -> per
sunposeLayer: instance-norm backward fused with the ReLU mask (also yields d gamma, d beta), distortion-aware conv weight and data
gradients (the 3-channel first layer on the small-C kernels).  All gradients land in ONE flat buffer: data-parallel training
all-reduces it once per step (NCCL) and the fused Adam kernel applies 1/world.  Needs the distortion-aware wiring of
sunpose_net.py:11,16 (the plain wiring has no weight-gradient kernel yet).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import grad_cam, tf_utils
from ._lib import LIB, check
from .distortion_aware_ops import _require_cuda, _stream, conv2d as da_conv2d, conv2d_backward
from .sunpose_net import maxpool2d_backward
from .trunk_train import allreduce_flat_

LEARNING_RATE = 1e-4       # train_sun.py:34


def start_tail_allreduce(flat_g, split):
    """Data-parallel gradient exchange, first half: asynchronously sum flat_g[split:] (the Dense gradients, complete early in the
    backward pass) over the ranks; returns the work handle, or None in a single-process run."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.all_reduce(flat_g[split:], op=dist.ReduceOp.SUM, async_op=True)
    return None


def finish_allreduce(flat_g, split, work):
    """Second half: sum flat_g[:split] (conv / norm gradients, complete at the end of the backward pass) and wait for the first."""
    if work is not None:
        allreduce_flat_(flat_g[:split])
        work.wait()


class SunTrainer:
    def __init__(self, net, batch_size, im_height=32, im_width=128, lr=LEARNING_RATE, beta1=0.9, beta2=0.999, eps=1e-7):
        self.net = net
        self.H, self.W = im_height, im_width
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.step_count = 0
        dev = net.fc1.device
        x0 = torch.zeros((batch_size, im_height, im_width, 3), dtype=torch.float32, device=dev)
        net.sunposeEstimation(x0)                                   # builds every lazily created variable
        for layer in (net.sunlayer1, net.sunlayer2, net.sunlayer3):
            if not isinstance(layer.conv1, da_conv2d):
                raise NotImplementedError("SunTrainer needs the distortion-aware wiring (sunpose_net.py:11,16)")
        # flat parameter / gradient / Adam-moment buffers; every variable becomes a view (order: trainable_variables)
        owners = []
        for layer in (net.sunlayer1, net.sunlayer2, net.sunlayer3):
            for conv, norm in ((layer.conv1, layer.norm1), (layer.conv2, layer.norm2)):
                owners += [(conv, "kernel"), (conv, "bias"), (norm, "gamma"), (norm, "beta")]
        owners += [(net.fc1, "kernel"), (net.fc1, "bias"), (net.fc2, "kernel"), (net.fc2, "bias")]
        shapes = [tuple(getattr(o, a).shape) for o, a in owners]
        sizes = [int(torch.tensor(s).prod()) for s in shapes]
        pad = [(-n) % 4 for n in sizes]
        total = sum(n + p for n, p in zip(sizes, pad))
        self.flat_w = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros_like(self.flat_w)
        self.flat_m = torch.zeros_like(self.flat_w)
        self.flat_v = torch.zeros_like(self.flat_w)
        self.grads, off = {}, 0
        self._fc_offset = None                                      # start of the Dense variables (99 % of the bytes) in the flat buffers
        for (o, a), n, p, s in zip(owners, sizes, pad, shapes):
            if o is net.fc1 and self._fc_offset is None:
                self._fc_offset = off
            vw = self.flat_w[off:off + n].view(s)
            vw.copy_(getattr(o, a))
            setattr(o, a, vw)                                       # the layer now reads the flat buffer
            self.grads[(id(o), a)] = self.flat_g[off:off + n].view(s)
            off += n + p
        self._side = None
        self._acc = torch.zeros(8, dtype=torch.float64, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float64, device=dev)

    def _g(self, obj, attr):
        return self.grads[(id(obj), attr)]

    def _layer_backward(self, layer, g_out, need_input_grad):
        x, conv1, actv1, conv2, actv2 = layer._saved
        B, h, w, F = conv2.shape
        sums = torch.empty(B, F, 2, dtype=torch.float64, device=x.device)

        def norm_bwd(norm, pre, stats, dy, act):
            dx = torch.empty_like(pre)
            check(LIB.sky_instnorm_bwd(pre.data_ptr(), stats.data_ptr(), norm.gamma.data_ptr(), dy.data_ptr(), act.data_ptr(), None,
                                       sums.data_ptr(), dx.data_ptr(), self._g(norm, "gamma").data_ptr(), self._g(norm, "beta").data_ptr(),
                                       B, h, w, F, norm.epsilon, 0.0, _stream()))
            return dx

        def conv_bwd(conv, xin, dy, need_dx):
            C = xin.shape[-1]
            dk, db = self._g(conv, "kernel"), self._g(conv, "bias")
            if C <= 4:                                              # sunlayer1.conv1 on the LDR panorama
                check(LIB.sky_da_conv2d_smallc_bwd_filter(xin.data_ptr(), dy.data_ptr(), conv.offset_table.data_ptr(), dk.data_ptr(),
                                                          db.data_ptr(), B, h, w, C, conv.filters, conv.kernel_size, _stream()))
                return None
            return conv2d_backward(conv, xin, dy, need_dx=need_dx, dk_out=dk, db_out=db)[0]

        g = norm_bwd(layer.norm2, conv2, layer._stats[1], g_out, actv2)          # relu (:28) + IN (:26)
        g = conv_bwd(layer.conv2, actv1, g, True)                                # conv2 (:25)
        g = norm_bwd(layer.norm1, conv1, layer._stats[0], g, actv1)              # relu (:23) + IN (:22)
        return conv_bwd(layer.conv1, x, g, need_input_grad)                      # conv1 (:21)

    def sun_train_step(self, ds, sunpose_gt):
        _, jpeg_img_float = ds                                                   # train_sun.py:223 (callers pass BGR, :225)
        x, gt = _require_cuda(jpeg_img_float, "jpeg_img_float"), _require_cuda(sunpose_gt, "sunpose_gt")
        net = self.net
        B = x.shape[0]
        sm, acts = net.sunposeEstimation(x, training=True)                       # :229
        # Grad-CAM (:231-240) is outside the tape: its backward sweep shares nothing with the loss backward but read-only forward
        # tensors, so it runs on a side stream next to it (most of its kernels fill a fraction of the SMs): 4.02 -> 3.72 ms per step.
        # (Putting the weight gradients on a further stream next to the data gradients was tried and is much slower: 6.05 ms.)
        net.fc1.kernel_transposed()                                              # both sweeps read the cached W^T: build them first
        net.fc2.kernel_transposed()
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            y_c = net.class_score(sm, gt)                                        # :234-236
            cams = [grad_cam.layer(y_c, a) for a in acts]                        # :238-240
        pred, sungt = sm.reshape(B, self.H, self.W, 1), gt.reshape(B, self.H, self.W, 1)   # :247-248
        # ---- loss (:245-255) ----
        acc = self._acc.zero_()
        check(LIB.sky_kl_divergence(gt.data_ptr(), sm.data_ptr(), gt.numel(), acc[0:1].data_ptr(), _stream()))
        base_p = torch.empty((B, 2 * self.H, 2 * self.W, 1), dtype=torch.float32, device=x.device)
        base_g = torch.empty_like(base_p)
        check(LIB.sky_dog_base(pred.data_ptr(), base_p.data_ptr(), B, self.H, self.W, 1, _stream()))
        check(LIB.sky_dog_base(sungt.data_ptr(), base_g.data_ptr(), B, self.H, self.W, 1, _stream()))
        check(LIB.sky_dog_l1(base_p.data_ptr(), base_g.data_ptr(), B, 2 * self.H, 2 * self.W, 1, acc[1:5].data_ptr(), _stream()))
        torch.add(acc[0:1] / B, acc[1:5].sum() / base_p.numel(), out=self.loss)
        # ---- backward (:257) ----
        self.flat_g[:self._fc_offset].zero_()                                     # d gamma / d beta accumulate with atomics (the Dense
                                                                                  # gradients behind them are overwritten, not accumulated)
        g_sm = torch.empty_like(sm)
        check(LIB.sky_kl_divergence_bwd(gt.data_ptr(), sm.data_ptr(), g_sm.data_ptr(), sm.numel(), 1.0 / B, 0, _stream()))
        dbase = torch.empty_like(base_p)
        check(LIB.sky_dog_l1_bwd(base_p.data_ptr(), base_g.data_ptr(), dbase.data_ptr(), B, 2 * self.H, 2 * self.W, 1,
                                 1.0 / base_p.numel(), _stream()))
        check(LIB.sky_dog_base_bwd(dbase.data_ptr(), g_sm.data_ptr(), B, self.H, self.W, 1, 1, _stream()))
        _, actv1_s, actv2_s, pool_shape, _ = net._saved
        flat = net._saved_flat
        g_z2 = torch.empty_like(sm)
        check(LIB.sky_softmax_bwd_rows(sm.data_ptr(), g_sm.data_ptr(), actv2_s.data_ptr(), g_z2.data_ptr(), B, sm.shape[1], _stream()))
        n_fc = sm.shape[1]
        check(LIB.sky_dense_bwd_filter(actv1_s.data_ptr(), g_z2.data_ptr(), self._g(net.fc2, "kernel").data_ptr(),
                                       self._g(net.fc2, "bias").data_ptr(), B, actv1_s.shape[1], n_fc, _stream()))
        g_z1 = net.fc2.backward_data(g_z2, act=actv1_s)
        check(LIB.sky_dense_bwd_filter(flat.data_ptr(), g_z1.data_ptr(), self._g(net.fc1, "kernel").data_ptr(),
                                       self._g(net.fc1, "bias").data_ptr(), B, flat.shape[1], n_fc, _stream()))
        # data-parallel: the Dense gradients (201 MB at 32x128) are complete here, before any conv gradient — their all-reduce is
        # started now and runs on NCCL's stream under the rest of the backward pass
        fc_work = start_tail_allreduce(self.flat_g, self._fc_offset)
        g = net.fc1.backward_data(g_z1)
        g = maxpool2d_backward(acts[2], g.view(pool_shape))
        g = self._layer_backward(net.sunlayer3, g, True)
        g = maxpool2d_backward(acts[1], g)
        g = self._layer_backward(net.sunlayer2, g, True)
        g = maxpool2d_backward(acts[0], g)
        self._layer_backward(net.sunlayer1, g, False)
        # ---- optimizer (:258) ----
        finish_allreduce(self.flat_g, self._fc_offset, fc_work)                   # conv / norm gradients (1 MB), then join
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        main.wait_stream(self._side)                                              # Grad-CAM done before Adam invalidates W^T / the packed kernels
        self.apply_gradients(world)
        return pred, sungt, cams

    def apply_gradients(self, world=1):
        self.step_count += 1
        check(LIB.sky_adam_step(self.flat_w.data_ptr(), self.flat_m.data_ptr(), self.flat_v.data_ptr(), self.flat_g.data_ptr(),
                                self.flat_w.numel(), self.lr, self.beta1, self.beta2, self.eps, self.step_count, 1.0 / world, _stream()))
        net = self.net
        for layer in (net.sunlayer1, net.sunlayer2, net.sunlayer3):                # packed TF32 copies / transposes are stale now
            layer.conv1._packed_key = None
            layer.conv2._packed_key = None
        net.fc1._kernel_t_key = None
        net.fc2._kernel_t_key = None
