"""GPU parity of the trunk train step against torch autograd through the fp64 oracle (== TF autodiff of generator.py:26-49
with the distortion-aware convs) and against the Keras RMSprop formula.
Tolerances: parameter / input gradients relative L2 <= 2e-2 (TF32 operands through 4 convs and 4 norms forward and
backward; the instance-norm backward subtracts two plane means, which amplifies the operand rounding);
one RMSprop step <= 1e-6 given the same gradients."""
import numpy as np
import pytest
import torch

from oracle import model_oracle as M

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def test_instnorm_backward_vs_autograd(pkg):
    torch.manual_seed(3)
    B, h, w, C = 2, 8, 32, 64
    x = torch.randn(B, h, w, C, dtype=torch.float64, requires_grad=True)
    gamma = (1 + 0.2 * torch.randn(C, dtype=torch.float64)).requires_grad_(True)
    beta = (0.2 * torch.randn(C, dtype=torch.float64)).requires_grad_(True)
    dy = torch.randn(B, h, w, C, dtype=torch.float64)
    extra = torch.randn(B, h, w, C, dtype=torch.float64)
    a = M.leaky_relu(M.instance_norm(x, gamma, beta), 0.1)
    a.backward(dy)
    xs = x.detach().float().cuda()
    stats = torch.stack([xs.double().sum((1, 2)), (xs.double() ** 2).sum((1, 2))], dim=-1).contiguous()
    sums = torch.zeros(B, C, 2, dtype=torch.float64, device="cuda")
    dx = torch.empty_like(xs)
    dg = torch.zeros(C, device="cuda")
    db = torch.zeros(C, device="cuda")
    L = pkg._lib
    gd, dyd, ad, ed = (t.detach().float().cuda() for t in (gamma, dy, a, extra))      # keep the device copies alive
    L.check(L.LIB.sky_instnorm_bwd(xs.data_ptr(), stats.data_ptr(), gd.data_ptr(), dyd.data_ptr(), ad.data_ptr(),
                                   ed.data_ptr(), sums.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(),
                                   B, h, w, C, 1e-3, 0.1, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert rel_l2(dx.cpu(), x.grad + extra) < 1e-5
    assert rel_l2(dg.cpu(), gamma.grad) < 1e-5 and rel_l2(db.cpu(), beta.grad) < 1e-5


def test_trunk_train_step_vs_oracle(pkg):
    rng = np.random.default_rng(1)
    B, h, w, C, k, nb = 2, 8, 32, 128, 3, 2
    x = rng.standard_normal((B, h, w, C)).astype(np.float32)
    tgt = rng.standard_normal((B, h, w, C)).astype(np.float32)
    blocks = M.random_trunk_weights(nb, C, k, seed=2)
    # oracle: fp64 autograd
    params = [{kk: v.double().clone().requires_grad_(True) for kk, v in b.items()} for b in blocks]
    xo = torch.from_numpy(x).double().requires_grad_(True)
    y = M.res_layer(xo, params, k, acc_dtype=torch.float64)
    loss = ((y - torch.from_numpy(tgt).double()) ** 2).mean()
    loss.backward()
    # ours
    trunk = pkg.resLayer((C,) * nb, C, k_h=k, k_w=k)
    trunk.build((B, h, w, C))
    trunk.set_weights(blocks)
    tr = pkg.trunk_train.TrunkTrainer(trunk, (B, h, w, C), lr=1e-3)
    xd, td = torch.from_numpy(x).cuda(), torch.from_numpy(tgt).cuda()
    yd = tr.forward(xd)
    l, dy = tr.loss_and_grad(yd, td)
    dx = tr.backward(dy)
    assert abs(l.item() - loss.item()) / loss.item() < 5e-3
    assert rel_l2(dx.cpu(), xo.grad) <= 2e-2, rel_l2(dx.cpu(), xo.grad)
    names = {"k": "kernel", "b": "bias", "g": "gamma", "be": "beta"}
    for g, p in zip(tr.blocks, params):
        for key, grad in g.items():
            short, idx = key[:-1], key[-1]
            ref = p[("conv" if short in ("k", "b") else "norm") + idx + "_" + names[short]].grad
            if short == "b":
                # a bias in front of an instance norm has exactly zero gradient (the norm removes the mean): both sides
                # must be numerically negligible next to the kernel gradient
                scale = float(g["k" + idx].norm())
                assert float(grad.norm()) <= 1e-3 * scale and float(ref.norm()) <= 1e-6 * scale, key
                continue
            assert rel_l2(grad.cpu(), ref) <= 2e-2, (key, rel_l2(grad.cpu(), ref))
    # one Keras-RMSprop step from these gradients
    w0, g0 = tr.flat_w.clone(), tr.flat_g.clone()
    tr.apply_gradients(world=1)
    ms = 0.1 * g0 ** 2
    want = w0 - 1e-3 * g0 / (ms.sqrt() + 1e-7)
    assert rel_l2((tr.flat_w - w0).cpu(), (want - w0).cpu()) < 1e-5
    # and the forward pass sees the updated weights (packed TF32 copies were invalidated)
    y2 = tr.forward(xd)
    assert not torch.equal(y2, yd)
