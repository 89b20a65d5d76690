"""Host-side mirror of the compute step of the reference's train_sun.py: ``sun_train_step`` (train_sun.py:220-264) — the pre-training
step of the sun-position network (BASELINE configs[1]):

    forward (sunposeEstimation) -> Grad-CAM maps at the ground-truth class (outside the tape, :231-240) -> loss = KLDivergence(gt, sm)
    + sum_l mean |DoG_l(pred) - DoG_l(gt)| (:245-255) -> gradients w.r.t. every variable of the network (:257) -> Adam (:258).

Backward: KL and DoG adjoints -> softmax backward (+ ReLU mask) -> Dense weight / data gradients -> max-pool gradient -> per
sunposeLayer: instance-norm backward fused with the ReLU mask (also yields d gamma, d beta), distortion-aware conv weight and data
gradients (the 3-channel first layer on the small-C kernels).  All gradients land in ONE flat buffer: data-parallel training
all-reduces it once per step (NCCL); the loss adjoints are normalised by the global batch, so the summed gradient is the global-batch
mean.  Both wirings of sunpose_net.py train: the distortion-aware one (:11,16) and the plain one that is live in the reference (:10,15).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import grad_cam, tf_utils
from ._lib import LIB, check
from .distortion_aware_ops import _require_cuda, _stream, conv2d as da_conv2d, conv2d_backward
from .sunpose_net import maxpool2d_backward
from .trunk_train import allreduce_flat_
from ._flat import FlatVars

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


def sunpose_owner_list(net):
    """(object, attribute) of every trainable variable of sunpose_net.model; conv / norm variables first, the Dense variables (99 % of
    the bytes) last, so that their gradients — the first ones the backward pass completes — form one contiguous all-reduce bucket."""
    owners = []
    for layer in (net.sunlayer1, net.sunlayer2, net.sunlayer3):
        for conv, norm in ((layer.conv1, layer.norm1), (layer.conv2, layer.norm2)):
            kname, bname = ("kernel", "bias") if isinstance(conv, da_conv2d) else ("w", "biases")
            owners += [(conv, kname), (conv, bname), (norm, "gamma"), (norm, "beta")]
    return owners, [(net.fc1, "kernel"), (net.fc1, "bias"), (net.fc2, "kernel"), (net.fc2, "bias")]


def sunpose_backward(net, g_sm, grad_of, on_dense_done=None, accumulate=False):
    """Backward of sunposeEstimation (sunpose_net.py:54-72) from the gradient w.r.t. its softmax output: softmax backward (+ ReLU mask),
    Dense weight / data gradients, max-pool gradients, per sunposeLayer the instance-norm backward fused with the ReLU mask and the conv
    weight / data gradients (either wiring; the 3-channel first layer on the small-C kernels).  grad_of(obj, attr) -> gradient view.
    on_dense_done() is called as soon as the Dense gradients are complete (start of their all-reduce bucket).  accumulate: the conv
    gradient views were zeroed by the caller (one memset of the flat buffer per step): the conv weight gradients add to them."""
    sm, actv1_s, actv2_s, pool_shape, acts = net._saved
    flat = net._saved_flat
    B, n_fc = sm.shape
    g_z2 = torch.empty_like(sm)
    check(LIB.sky_softmax_bwd_rows(sm.data_ptr(), g_sm.data_ptr(), actv2_s.data_ptr(), g_z2.data_ptr(), B, n_fc, _stream()))
    check(LIB.sky_dense_bwd_filter(actv1_s.data_ptr(), g_z2.data_ptr(), grad_of(net.fc2, "kernel").data_ptr(),
                                   grad_of(net.fc2, "bias").data_ptr(), B, actv1_s.shape[1], n_fc, _stream()))
    g_z1 = net.fc2.backward_data(g_z2, act=actv1_s)
    check(LIB.sky_dense_bwd_filter(flat.data_ptr(), g_z1.data_ptr(), grad_of(net.fc1, "kernel").data_ptr(),
                                   grad_of(net.fc1, "bias").data_ptr(), B, flat.shape[1], n_fc, _stream()))
    if on_dense_done is not None:
        on_dense_done()
    g = net.fc1.backward_data(g_z1)
    g = maxpool2d_backward(acts[2], g.view(pool_shape))

    def layer_backward(layer, g_out, need_input_grad):
        x, conv1, actv1, conv2, actv2 = layer._saved
        Bq, h, w, F = conv2.shape
        sums = torch.empty(Bq, F, 2, dtype=torch.float64, device=x.device)

        def norm_bwd(norm, pre, stats, dy, act):
            dx = torch.empty_like(pre)
            check(LIB.sky_instnorm_bwd(pre.data_ptr(), stats.data_ptr(), norm.gamma.data_ptr(), dy.data_ptr(), act.data_ptr(), None,
                                       sums.data_ptr(), dx.data_ptr(), grad_of(norm, "gamma").data_ptr(), grad_of(norm, "beta").data_ptr(),
                                       Bq, h, w, F, norm.epsilon, 0.0, _stream()))
            return dx

        def conv_bwd(conv, xin, dy, need_dx):
            C = xin.shape[-1]
            if isinstance(conv, da_conv2d):
                dk, db = grad_of(conv, "kernel"), grad_of(conv, "bias")
                if C <= 4:                                          # sunlayer1.conv1 on the LDR panorama
                    check(LIB.sky_da_conv2d_smallc_bwd_filter(xin.data_ptr(), dy.data_ptr(), conv.offset_table.data_ptr(), dk.data_ptr(),
                                                              db.data_ptr(), Bq, h, w, C, conv.filters, conv.kernel_size, _stream()))
                    return None
                return conv2d_backward(conv, xin, dy, need_dx=need_dx, dk_out=dk, db_out=db, accumulate_dw=accumulate)[0]
            conv.backward_filter(xin, dy, grad_of(conv, "w"), grad_of(conv, "biases"), accumulate=accumulate and C > 4)   # plain wiring (sunpose_net.py:10,15)
            return conv.backward_data(xin, dy) if (need_dx and C > 4) else None

        g_ = norm_bwd(layer.norm2, conv2, layer._stats[1], g_out, actv2)          # relu (:28) + IN (:26)
        g_ = conv_bwd(layer.conv2, actv1, g_, True)                               # conv2 (:25)
        g_ = norm_bwd(layer.norm1, conv1, layer._stats[0], g_, actv1)             # relu (:23) + IN (:22)
        return conv_bwd(layer.conv1, x, g_, need_input_grad)                      # conv1 (:21)

    g = layer_backward(net.sunlayer3, g, True)
    g = maxpool2d_backward(acts[1], g)
    g = layer_backward(net.sunlayer2, g, True)
    g = maxpool2d_backward(acts[0], g)
    layer_backward(net.sunlayer1, g, False)


class SunTrainer:
    def __init__(self, net, batch_size, im_height=32, im_width=128, lr=LEARNING_RATE, beta1=0.9, beta2=0.999, eps=1e-7):
        self.net = net
        self.H, self.W = im_height, im_width
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.step_count = 0
        dev = net.fc1.device
        x0 = torch.zeros((max(int(batch_size), 1), im_height, im_width, 3), dtype=torch.float32, device=dev)
        net.sunposeEstimation(x0)                                   # builds every lazily created variable
        # flat parameter / gradient / Adam-moment buffers; every variable becomes a view (conv / norm first, Dense last)
        conv_owners, fc_owners = sunpose_owner_list(net)
        self.fv = FlatVars(conv_owners + fc_owners, dev)
        self.flat_w, self.flat_g = self.fv.flat_w, self.fv.flat_g
        self.flat_m = torch.zeros_like(self.flat_w)
        self.flat_v = torch.zeros_like(self.flat_w)
        self._fc_offset = self.fv.offset(net.fc1, "kernel")          # start of the Dense variables (99 % of the bytes)
        self._side = None
        self._acc = torch.zeros(8, dtype=torch.float64, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float64, device=dev)

    def _g(self, obj, attr):
        return self.fv.grad(obj, attr)

    def sun_train_step(self, ds, sunpose_gt, global_batch=None):
        """One step on this rank's shard.  Data-parallel runs normalise the loss adjoints by the GLOBAL batch (`global_batch`, default
        local batch x world size), so the all-reduced sum of the per-rank gradients is the gradient of the reference's global-batch
        mean whatever the shard sizes; a rank with an empty shard skips the compute and joins both all-reduces with zeros."""
        _, jpeg_img_float = ds                                                   # train_sun.py:223 (callers pass BGR, :225)
        x, gt = _require_cuda(jpeg_img_float, "jpeg_img_float"), _require_cuda(sunpose_gt, "sunpose_gt")
        net = self.net
        B = x.shape[0]
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        Bg = int(global_batch) if global_batch is not None else B * world
        main = torch.cuda.current_stream()
        if B == 0:
            self.flat_g.zero_()
            self.loss.zero_()
            finish_allreduce(self.flat_g, self._fc_offset, start_tail_allreduce(self.flat_g, self._fc_offset))
            self.apply_gradients()
            return None, None, []
        sm, acts = net.sunposeEstimation(x, training=True)                       # :229
        # Grad-CAM (:231-240) is outside the tape: its backward sweep shares nothing with the loss backward but read-only forward
        # tensors, so it runs on a side stream next to it (most of its kernels fill a fraction of the SMs): 4.02 -> 3.72 ms per step.
        # (Putting the weight gradients on a further stream next to the data gradients was tried and is much slower: 6.05 ms.)
        net.fc1.kernel_transposed()                                              # both sweeps read the cached W^T: build them first
        net.fc2.kernel_transposed()
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
        check(LIB.sky_kl_divergence_bwd(gt.data_ptr(), sm.data_ptr(), g_sm.data_ptr(), sm.numel(), 1.0 / Bg, 0, _stream()))
        dbase = torch.empty_like(base_p)
        check(LIB.sky_dog_l1_bwd(base_p.data_ptr(), base_g.data_ptr(), dbase.data_ptr(), B, 2 * self.H, 2 * self.W, 1,
                                 1.0 / (base_p.numel() // B * Bg), _stream()))
        check(LIB.sky_dog_base_bwd(dbase.data_ptr(), g_sm.data_ptr(), B, self.H, self.W, 1, 1, _stream()))
        # data-parallel: the Dense gradients (201 MB at 32x128) are complete before any conv gradient — their all-reduce is started
        # from the callback and runs on NCCL's stream under the rest of the backward pass
        work = []
        sunpose_backward(net, g_sm, self._g, on_dense_done=lambda: work.append(start_tail_allreduce(self.flat_g, self._fc_offset)), accumulate=True)
        # ---- optimizer (:258) ----
        finish_allreduce(self.flat_g, self._fc_offset, work[0])                   # conv / norm gradients (1 MB), then join
        main.wait_stream(self._side)                                              # Grad-CAM done before Adam invalidates W^T / the packed kernels
        self.apply_gradients()
        return pred, sungt, cams

    def apply_gradients(self, world=1):
        """Keras Adam over the flat buffer.  The gradients are already normalised by the global batch: grad_scale stays 1."""
        self.step_count += 1
        check(LIB.sky_adam_step(self.flat_w.data_ptr(), self.flat_m.data_ptr(), self.flat_v.data_ptr(), self.flat_g.data_ptr(),
                                self.flat_w.numel(), self.lr, self.beta1, self.beta2, self.eps, self.step_count, 1.0, _stream()))
        self.fv.invalidate()                                                      # packed TF32 copies / transposes are stale now
