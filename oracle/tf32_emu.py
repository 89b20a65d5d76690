"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/da_oracle.py for the rules).

TF32 emulation for the oracle: what the reference's TensorFlow path computes on any Ampere-or-newer GPU, where
``tf.config.experimental.enable_tensor_float_32_execution`` defaults to True (TensorFlow >= 2.4, the version README.md:25 asks
for): tf.matmul (distortion_aware_ops.py:117) and tf.nn.conv2d / Keras Conv2D (ops.py:41, sunrad_net.py:12) round both operands
to 10 mantissa bits and accumulate in fp32.  Inside ``with emulate():`` the oracle's contractions round their operands the same
way (round-to-nearest, ties away from zero — PTX cvt.rna.tf32.f32, the conversion the sm_100a kernels use) and accumulate in the
caller's dtype (fp64 in the tests), forward AND backward: the data gradient contracts round(dy) with round(W), the weight
gradient round(x) with round(dy), like the tensor-core backward kernels.  Comparing the GPU's `tf32` mode against this oracle
separates "the kernels are right given TF32 operands" from "TF32 is coarser than fp32".
"""
from __future__ import annotations

import contextlib

import torch

ENABLED = False


@contextlib.contextmanager
def emulate(on=True):
    global ENABLED
    prev, ENABLED = ENABLED, bool(on)
    try:
        yield
    finally:
        ENABLED = prev


def round_tf32(t: torch.Tensor) -> torch.Tensor:
    """fp32 (or fp64, first rounded to fp32) -> nearest TF32 value, ties away from zero, returned in t's dtype."""
    f = t.detach().to(torch.float32).contiguous()
    bits = f.view(torch.int32)
    bits = (bits + 0x1000) & ~0x1FFF          # sign-magnitude: adding to the low 31 bits grows the magnitude for either sign
    out = bits.view(torch.float32)
    out = torch.where(torch.isfinite(f), out, f)
    return out.to(t.dtype)


class _MatmulTF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ar, br = round_tf32(a), round_tf32(b)
        ctx.save_for_backward(ar, br)
        return ar @ br

    @staticmethod
    def backward(ctx, dy):
        ar, br = ctx.saved_tensors
        dyr = round_tf32(dy)
        da = dyr @ br.transpose(-1, -2)
        db = ar.transpose(-1, -2) @ dyr
        while db.dim() > br.dim():
            db = db.sum(0)
        return da, db


class _Conv2dTF32(torch.autograd.Function):
    """NCHW input (already padded), OIHW weight, no bias."""

    @staticmethod
    def forward(ctx, xp, w, stride):
        xr, wr = round_tf32(xp), round_tf32(w)
        ctx.save_for_backward(xr, wr)
        ctx.stride = stride
        return torch.nn.functional.conv2d(xr, wr, stride=stride)

    @staticmethod
    def backward(ctx, dy):
        xr, wr = ctx.saved_tensors
        dyr = round_tf32(dy)
        dx = torch.nn.grad.conv2d_input(xr.shape, wr, dyr, stride=ctx.stride)
        dw = torch.nn.grad.conv2d_weight(xr, wr.shape, dyr, stride=ctx.stride)
        return dx, dw, None


def matmul(a, b, tf32=True):
    if ENABLED and tf32:
        return _MatmulTF32.apply(a, b)
    return torch.matmul(a, b)


def conv2d(xp, w, stride=1, tf32=True):
    """tf32 = True: both operands rounded, forward and backward; "weights": the small-filter kernel of the library (3-filter layers
    conv1_f / conv1_u) multiplies fp32 activations with the TF32-rounded packed kernel on the fp32 pipes — forward only, its backward
    passes use the unrounded variable."""
    if ENABLED and tf32 == "weights":
        return torch.nn.functional.conv2d(xp.contiguous(), (w + (round_tf32(w) - w).detach()).contiguous(), stride=stride)
    if ENABLED and tf32:
        return _Conv2dTF32.apply(xp, w, stride)
    return torch.nn.functional.conv2d(xp.contiguous(), w.contiguous(), stride=stride)
