"""The weight gradient on the row-strip formulation (csrc/strip_wgrad.cu) against the per-tap gather kernels of round 1 and against
autograd through the oracle: distortion-aware layers (effective-weight plans: a window's accumulator is added to its taps with the
bilinear factors) and plain SAME convolutions (stride 1 / 2, even kernels, ragged batches and widths, more than 128 channels / filters,
3 filters).  Tolerance: TF32 operands, fp32 accumulate -> relative L2 <= 2e-3 against fp64; strip vs gather kernel <= 1e-3."""
import os
import unittest.mock as mock

import numpy as np
import pytest
import torch

from oracle import da_oracle as O
from oracle import model_oracle as M

pytestmark = pytest.mark.gpu
T = torch.from_numpy


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


DA_CASES = [  # (B, h, w, C, F, k)
    (2, 8, 32, 128, 128, 3),     # trunk
    (9, 8, 32, 128, 128, 3),     # two panorama tiles, the second ragged
    (3, 5, 24, 64, 32, 3),       # 64 channels (half of the MMA's lanes unused), odd height
    (2, 16, 64, 32, 32, 7),      # 32 channels: four windows per MMA; 7x7 strips split by span
    (2, 32, 128, 32, 32, 7),     # the sun-position net's second layer: one exact-tap strip at this geometry
    (2, 8, 20, 32, 64, 3),       # width not a multiple of the tile
    (1, 8, 32, 64, 128, 3),
    (2, 12, 48, 32, 96, 5),
]


@pytest.mark.parametrize("B,h,w,C,F,k", DA_CASES)
def test_da_strip_wgrad(pkg, B, h, w, C, F, k):
    rng = np.random.default_rng(B * 7 + h + C + F + k)
    x = rng.standard_normal((B, h, w, C)).astype(np.float32)
    kern = (rng.standard_normal((k * k * C, F)) / np.sqrt(k * k * C)).astype(np.float32)
    bias = np.zeros(F, np.float32)
    dy = rng.standard_normal((B, h, w, F)).astype(np.float32)
    _, want_dk, want_db = (g.numpy() for g in O.conv2d_backward(x, kern, bias, dy, k, acc_dtype=torch.float64))
    layer = pkg.conv2d(F, kernel_size=k, kernel_initializer=kern, bias_initializer=bias)
    layer.build(x.shape)
    D = pkg.distortion_aware_ops
    assert D.DA_WGRAD_KERNEL == "strip"
    xd, dyd = T(x).cuda(), T(dy).cuda()
    _, dk, db = D.conv2d_backward(layer, xd, dyd, need_dx=False)
    with mock.patch.object(D, "DA_WGRAD_KERNEL", "gather"):
        _, dk_g, db_g = D.conv2d_backward(layer, xd, dyd, need_dx=False)
    assert rel(dk.cpu().numpy(), want_dk) < 2e-3, rel(dk.cpu().numpy(), want_dk)
    assert rel(dk.cpu().numpy(), dk_g.cpu().numpy()) < 1e-3, rel(dk.cpu().numpy(), dk_g.cpu().numpy())
    assert rel(db.cpu().numpy(), want_db) < 1e-5


PLAIN_CASES = [  # (B, h, w, C, F, k, stride)
    (2, 16, 32, 32, 64, 3, 2),
    (9, 8, 16, 64, 128, 4, 2),
    (2, 4, 16, 256, 512, 4, 1),
    (3, 4, 16, 512, 1, 4, 1),
    (2, 16, 32, 32, 3, 7, 1),
    (2, 16, 40, 64, 32, 3, 1),
    (2, 8, 16, 128, 256, 3, 1),
    (2, 9, 15, 32, 64, 3, 2),
    (2, 7, 9, 32, 32, 1, 1),
    (2, 8, 16, 96, 48, 3, 1),
]


@pytest.mark.parametrize("case", PLAIN_CASES, ids=lambda c: "B%d_%dx%d_C%d_F%d_k%d_s%d" % c)
def test_plain_strip_wgrad(pkg, case):
    B, h, w, C, F, k, s = case
    rng = np.random.default_rng(hash(case) % 2 ** 31)
    x = rng.standard_normal((B, h, w, C)).astype(np.float32)
    w4 = (rng.standard_normal((k, k, C, F)) / np.sqrt(k * k * C)).astype(np.float32)
    xt, wt = T(x).double(), T(w4).double().requires_grad_(True)
    y = M.conv2d_same(xt, wt, torch.zeros(F, dtype=torch.float64), stride=s, acc_dtype=torch.float64)
    dy = rng.standard_normal(tuple(y.shape)).astype(np.float32)
    y.backward(T(dy).double())
    want = wt.grad.numpy().reshape(k * k * C, F)
    xd, dyd = T(x).cuda(), T(dy).cuda()
    dk = torch.full((k * k * C, F), 7.0, device="cuda")
    pkg.ops.conv_backward_filter(xd, dyd, k, s, dk, None)
    assert rel(dk.cpu().numpy(), want) < 2e-3, rel(dk.cpu().numpy(), want)
    os.environ["SKY_WGRAD_KERNEL"] = "gather"
    try:
        dk_g = torch.empty_like(dk)
        pkg.ops.conv_backward_filter(xd, dyd, k, s, dk_g, None)
    finally:
        del os.environ["SKY_WGRAD_KERNEL"]
    assert rel(dk.cpu().numpy(), dk_g.cpu().numpy()) < 1e-3, rel(dk.cpu().numpy(), dk_g.cpu().numpy())
    pkg.ops.conv_backward_filter(xd, dyd, k, s, dk, None, accumulate=True)
    assert rel(dk.cpu().numpy(), 2 * want) < 2e-3


def test_strip_wgrad_full_batch_vs_gather(pkg):
    """The layer shapes of the train step at the real batch (B = 32): strip vs gather kernel (no CPU oracle at this size)."""
    torch.manual_seed(3)
    D = pkg.distortion_aware_ops
    for (h, w, C, F, k) in ((8, 32, 128, 128, 3), (32, 128, 32, 32, 7), (16, 64, 64, 64, 3)):
        x, dy = torch.randn(32, h, w, C, device="cuda"), torch.randn(32, h, w, F, device="cuda")
        layer = pkg.conv2d(F, kernel_size=k)
        layer.build(tuple(x.shape))
        _, dk, _ = D.conv2d_backward(layer, x, dy, need_dx=False)
        with mock.patch.object(D, "DA_WGRAD_KERNEL", "gather"):
            _, dk_g, _ = D.conv2d_backward(layer, x, dy, need_dx=False)
        assert rel(dk.cpu().numpy(), dk_g.cpu().numpy()) < 1e-3, ((h, w, C, F, k), rel(dk.cpu().numpy(), dk_g.cpu().numpy()))
    for (h, w, C, F, k, s) in ((32, 128, 64, 32, 3, 1), (16, 64, 128, 64, 3, 1), (4, 16, 256, 512, 4, 1), (32, 128, 32, 3, 7, 1), (16, 64, 64, 128, 4, 2)):
        x = torch.randn(32, h, w, C, device="cuda")
        dy = torch.randn(32, (h + s - 1) // s, (w + s - 1) // s, F, device="cuda")
        dk, dk_g = torch.empty(k * k * C, F, device="cuda"), torch.empty(k * k * C, F, device="cuda")
        pkg.ops.conv_backward_filter(x, dy, k, s, dk, None)
        os.environ["SKY_WGRAD_KERNEL"] = "gather"
        try:
            pkg.ops.conv_backward_filter(x, dy, k, s, dk_g, None)
        finally:
            del os.environ["SKY_WGRAD_KERNEL"]
        assert rel(dk.cpu().numpy(), dk_g.cpu().numpy()) < 1e-3, ((h, w, C, F, k, s), rel(dk.cpu().numpy(), dk_g.cpu().numpy()))
