"""Data-parallel train step of the distortion-aware residual trunk (generator.py:9-49 with :14,18 wired in).

What a step does (one process per GPU):
    forward (saving conv outputs / moments / activations) -> L2 objective -> backward through every res-block
    (instance-norm backward fused with the LeakyReLU mask, distortion-aware dgrad / wgrad / dbias) into ONE flat gradient
    buffer -> a single NCCL all-reduce of that buffer -> one fused Keras-RMSprop launch over the flat parameter buffer
    (1/world folded into the update).
The reference trains with Keras RMSprop(lr=1e-4) on generator + sun variables (train.py:201-202, 402-403); this module is
that update restricted to the trunk, with a synthetic regression target instead of the GAN losses (not built yet).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib
from ._lib import LIB, check
from .distortion_aware_ops import _stream, conv2d_backward
from .generator import resLayer


def allreduce_flat_(flat: torch.Tensor):
    """The step's single collective: in-place sum over ranks of the flat gradient buffer (NCCL on GPUs, gloo in tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


class TrunkTrainer:
    def __init__(self, trunk: resLayer, input_shape, lr=1e-4, rho=0.9, eps=1e-7):
        self.trunk = trunk
        trunk.build(tuple(input_shape))
        self.lr, self.rho, self.eps = lr, rho, eps
        dev = trunk.sequence[0].conv1.kernel.device
        # flat parameter / gradient / RMSprop-state buffers; every variable becomes a view (order: trainable_variables)
        shapes = [tuple(v.shape) for v in trunk.trainable_variables]
        sizes = [int(torch.tensor(s).prod()) for s in shapes]
        pad = [(-n) % 4 for n in sizes]                       # keep every view 16-byte aligned
        total = sum(n + p for n, p in zip(sizes, pad))
        self.flat_w = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros_like(self.flat_w)
        self.flat_ms = torch.zeros_like(self.flat_w)
        self._views_w, self._views_g, off = [], [], 0
        for n, p, s in zip(sizes, pad, shapes):
            self._views_w.append(self.flat_w[off:off + n].view(s))
            self._views_g.append(self.flat_g[off:off + n].view(s))
            off += n + p
        it = iter(zip(self._views_w, self._views_g))
        self.blocks = []
        for unit in trunk.sequence:
            g = {}
            for conv, norm, tag in ((unit.conv1, unit.norm1, "1"), (unit.conv2, unit.norm2, "2")):
                for obj, attr, name in ((conv, "kernel", "k"), (conv, "bias", "b"), (norm, "gamma", "g"), (norm, "beta", "be")):
                    vw, vg = next(it)
                    vw.copy_(getattr(obj, attr))
                    setattr(obj, attr, vw)                     # the layer now reads the flat buffer
                    g[name + tag] = vg
            self.blocks.append(g)
        B, h, w, C = input_shape
        self._saved = None
        self._sums = torch.zeros(B, C, 2, dtype=torch.float64, device=dev)
        self._loss = torch.zeros(1, dtype=torch.float64, device=dev)
        self._stats = [torch.zeros(2, B, C, 2, dtype=torch.float64, device=dev) for _ in trunk.sequence]

    # ---- forward with saved tensors (resBlock.call, generator.py:26-35) ------------------------------------------------------
    def forward(self, x):
        saved = []
        for unit, st in zip(self.trunk.sequence, self._stats):
            st.zero_()
            c1 = unit.conv1.call(x, stats=st[0])
            a1 = unit.norm1.apply(c1, st[0], leaky_slope=0.1)
            c2 = unit.conv2.call(a1, stats=st[1])
            out = unit.norm2.apply(c2, st[1], residual=x)
            saved.append((x, c1, a1, c2))
            x = out
        self._saved = saved
        return x

    def loss_and_grad(self, y, target):
        dy = torch.empty_like(y)
        check(LIB.sky_mse_loss(y.data_ptr(), target.data_ptr(), dy.data_ptr(), self._loss.data_ptr(), y.numel(), _stream()))
        return self._loss, dy

    # ---- backward --------------------------------------------------------------------------------------------------------------
    def _in_bwd(self, x, stats, norm, dy, act, dgamma, dbeta):
        B, h, w, F = x.shape
        dx = torch.empty_like(x)
        check(LIB.sky_instnorm_bwd(x.data_ptr(), stats.data_ptr(), norm.gamma.data_ptr(), dy.data_ptr(),
                                   None if act is None else act.data_ptr(), None, self._sums.data_ptr(), dx.data_ptr(),
                                   dgamma.data_ptr(), dbeta.data_ptr(), B, h, w, F, norm.epsilon, 0.1, _stream()))
        return dx

    def backward(self, dout):
        self.flat_g.zero_()                                    # dgamma / dbeta accumulate with atomics
        for unit, st, g, (x, c1, a1, c2) in reversed(list(zip(self.trunk.sequence, self._stats, self.blocks, self._saved))):
            dc2 = self._in_bwd(c2, st[1], unit.norm2, dout, None, g["g2"], g["be2"])               # IN2 (generator.py:33)
            da1, _, _ = conv2d_backward(unit.conv2, a1, dc2, dk_out=g["k2"], db_out=g["b2"])        # conv2 (:32)
            dc1 = self._in_bwd(c1, st[0], unit.norm1, da1, a1, g["g1"], g["be1"])                  # lrelu (:30) + IN1 (:29)
            dx = dout.clone()                                                                        # identity branch (:35)
            conv2d_backward(unit.conv1, x, dc1, dx_out=dx, dk_out=g["k1"], db_out=g["b1"], accumulate_dx=True)   # conv1 (:28)
            dout = dx
        return dout

    # ---- optimizer --------------------------------------------------------------------------------------------------------------
    def apply_gradients(self, world=1):
        check(LIB.sky_rmsprop_step(self.flat_w.data_ptr(), self.flat_ms.data_ptr(), self.flat_g.data_ptr(), self.flat_w.numel(),
                                   self.lr, self.rho, self.eps, 1.0 / world, _stream()))
        for unit in self.trunk.sequence:                       # the packed TF32 copies are stale now
            unit.conv1._packed_key = None
            unit.conv2._packed_key = None

    def train_step(self, x, target):
        y = self.forward(x)
        loss, dy = self.loss_and_grad(y, target)
        self.backward(dy)
        allreduce_flat_(self.flat_g)                           # the step's single collective
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.apply_gradients(world)
        return loss
