"""GPU parity of the distortion-aware conv backward (dgrad, wgrad, dbias) against autograd through the oracle's
materialised forward (== what TF autodiff derives from distortion_aware_ops.py:50-123).
Tolerance: TF32 operands, fp32 accumulate -> relative L2 <= 2e-3 against the fp64 oracle."""
import numpy as np
import pytest
import torch

from oracle import da_oracle as O

pytestmark = pytest.mark.gpu
TOL = 2e-3


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


CASES = [
    (2, 8, 32, 128, 128, 3),   # res-trunk site
    (1, 8, 32, 32, 64, 3),     # C != F
    (3, 5, 24, 64, 32, 3),     # ragged: M = 360, odd h
    (1, 16, 64, 32, 32, 7),    # 7x7: 49 taps -> 13 tap groups, last one partial
    (2, 12, 48, 32, 96, 5),    # k = 5, F = 96
]


@pytest.mark.parametrize("B,h,w,C,F,k", CASES)
def test_backward_vs_oracle(pkg, B, h, w, C, F, k):
    rng = np.random.default_rng(B + h + C + F + k)
    x = rng.standard_normal((B, h, w, C)).astype(np.float32)
    kern = (rng.standard_normal((k * k * C, F)) / np.sqrt(k * k * C)).astype(np.float32)
    b = rng.standard_normal(F).astype(np.float32)
    dy = rng.standard_normal((B, h, w, F)).astype(np.float32)
    want_dx, want_dk, want_db = (g.numpy() for g in O.conv2d_backward(x, kern, b, dy, k, acc_dtype=torch.float64))
    layer = pkg.conv2d(F, kernel_size=k, kernel_initializer=kern, bias_initializer=b)
    layer.build(x.shape)
    dx, dk, db = pkg.distortion_aware_ops.conv2d_backward(layer, torch.from_numpy(x).cuda(), torch.from_numpy(dy).cuda())
    assert rel_l2(dx.cpu().numpy(), want_dx) <= TOL, ("dx", rel_l2(dx.cpu().numpy(), want_dx))
    # the default data gradient is the row-strip kernel over the transposed plan (gather form, no atomics); the scatter kernel of round 1
    # must agree, and `accumulate` adds to what dx holds
    D = pkg.distortion_aware_ops
    assert D.DA_BACKWARD_KERNEL == "strip" and layer._strip_t is not None
    import unittest.mock as mock
    with mock.patch.object(D, "DA_BACKWARD_KERNEL", "scatter"):
        dx_s = D.conv2d_backward(layer, torch.from_numpy(x).cuda(), torch.from_numpy(dy).cuda(), need_dw=False)[0]
    assert rel_l2(dx_s.cpu().numpy(), want_dx) <= TOL, ("dx scatter", rel_l2(dx_s.cpu().numpy(), want_dx))
    base = torch.from_numpy(rng.standard_normal(x.shape).astype(np.float32)).cuda()
    acc = base.clone()
    D.conv2d_backward(layer, torch.from_numpy(x).cuda(), torch.from_numpy(dy).cuda(), need_dw=False, dx_out=acc, accumulate_dx=True)
    assert rel_l2((acc - base).cpu().numpy(), want_dx) <= TOL
    assert rel_l2(dk.cpu().numpy(), want_dk) <= TOL, ("dkernel", rel_l2(dk.cpu().numpy(), want_dk))
    assert rel_l2(db.cpu().numpy(), want_db) <= 1e-5, ("dbias", rel_l2(db.cpu().numpy(), want_db))


def test_backward_adjoint_identity_full_size(pkg):
    """<conv(x), dy> == <x, dgrad(dy)> + <bias-free part>: the adjoint identity needs no CPU oracle, so it runs at a
    BASELINE config-4 size (B=64, 32x128, 128->128)."""
    torch.manual_seed(0)
    B, h, w, C, F, k = 64, 32, 128, 128, 128, 3
    layer = pkg.conv2d(F, kernel_size=k, math_mode="3xtf32")
    x = torch.randn(B, h, w, C, device="cuda")
    dy = torch.randn(B, h, w, F, device="cuda")
    layer.build(tuple(x.shape))
    y = layer(x)                                    # bias is zero-initialised
    dx, dk, db = pkg.distortion_aware_ops.conv2d_backward(layer, x, dy)
    lhs = (y.double() * dy.double()).sum().item()
    rhs_x = (x.double() * dx.double()).sum().item()
    rhs_w = (layer.kernel.double() * dk.double()).sum().item()
    scale = (y.double().norm() * dy.double().norm()).item()
    assert abs(lhs - rhs_x) / scale < 2e-3
    assert abs(lhs - rhs_w) / scale < 2e-3
    assert torch.allclose(db, dy.sum((0, 1, 2)), rtol=1e-3, atol=1e-1)
