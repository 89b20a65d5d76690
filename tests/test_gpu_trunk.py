"""GPU parity of the residual trunk (generator.py:9-49 with the distortion-aware convs of :14,18) against the oracle.

Tolerances (relative L2 against the fp64-accumulated oracle on identical weights):
  one res-block, TF32      <= 2e-3          six-block trunk, TF32    <= 4e-3
  one res-block, 3xTF32    <= 5e-5          six-block trunk, 3xTF32  <= 1e-4
"""
import numpy as np
import pytest
import torch

from oracle import model_oracle as M

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def test_instnorm_apply_from_moments(pkg):
    torch.manual_seed(1)
    B, h, w, C = 3, 8, 32, 64
    x = torch.randn(B, h, w, C, device="cuda") * 3 + 1.5
    res = torch.randn_like(x)
    norm = pkg.InstanceNormalization()
    norm.build(tuple(x.shape))
    norm.gamma.normal_(1.0, 0.2)
    norm.beta.normal_(0.0, 0.2)
    stats = torch.stack([x.double().sum((1, 2)), (x.double() ** 2).sum((1, 2))], dim=-1).contiguous()
    want = M.instance_norm(x.double().cpu(), norm.gamma.double().cpu(), norm.beta.double().cpu())
    got = norm.apply(x, stats).cpu()
    assert rel_l2(got, want) < 2e-6
    got = norm.apply(x, stats, leaky_slope=0.1).cpu()
    assert rel_l2(got, M.leaky_relu(want, 0.1)) < 2e-6
    got = norm.apply(x, stats, residual=res).cpu()
    assert rel_l2(got, want + res.double().cpu()) < 2e-6


@pytest.mark.parametrize("mode,tol_block,tol_trunk", [("tf32", 2e-3, 4e-3), ("3xtf32", 5e-5, 1e-4)])
def test_res_block_and_trunk_vs_oracle(pkg, mode, tol_block, tol_trunk):
    rng = np.random.default_rng(5)
    B, h, w, C, k = 2, 8, 32, 128, 3          # the trunk site for 32x128 panoramas
    x = rng.standard_normal((B, h, w, C)).astype(np.float32)
    blocks = M.random_trunk_weights(6, C, k, seed=9)
    xd = torch.from_numpy(x).cuda()

    block = pkg.resBlock(C, C, k_h=k, k_w=k, math_mode=mode)
    block.build(tuple(xd.shape))
    block.set_weights(blocks[0])
    want = M.res_block(torch.from_numpy(x), blocks[0], k, acc_dtype=torch.float64).numpy()
    got = block(xd).cpu().numpy()
    assert rel_l2(got, want) <= tol_block, rel_l2(got, want)

    trunk = pkg.resLayer((C,) * 6, C, k_h=k, k_w=k, math_mode=mode)
    trunk.build(tuple(xd.shape))
    trunk.set_weights(blocks)
    want = M.res_layer(x, blocks, k, acc_dtype=torch.float64).numpy()
    got = trunk(xd).cpu().numpy()
    assert rel_l2(got, want) <= tol_trunk, rel_l2(got, want)
    # the fp32 oracle (what a CPU TensorFlow run would compute) sits inside the same band
    want32 = M.res_layer(x, blocks, k, acc_dtype=torch.float32).numpy()
    assert rel_l2(got, want32) <= tol_trunk


def test_projection_shortcut(pkg):
    """resBlock with filter_in != filter_out (generator.py:21-24): the identity branch is a 1x1 ops.conv2d.  Forward and the gradient
    w.r.t. the block input / the projection kernel against autograd through the oracle (TF32 operands: 2e-3 / 2e-2)."""
    import numpy as np
    from oracle import model_oracle as M
    rng = np.random.default_rng(9)
    B, h, w, Ci, Co, k = 2, 8, 32, 64, 128, 3
    x = rng.standard_normal((B, h, w, Ci)).astype(np.float32)
    wts = {}
    c = Ci
    for i in (1, 2):
        wts[f"conv{i}_kernel"] = (rng.standard_normal((k * k * c, Co)) / np.sqrt(k * k * c)).astype(np.float32)
        wts[f"conv{i}_bias"] = (0.1 * rng.standard_normal(Co)).astype(np.float32)
        wts[f"norm{i}_gamma"] = (1 + 0.1 * rng.standard_normal(Co)).astype(np.float32)
        wts[f"norm{i}_beta"] = (0.1 * rng.standard_normal(Co)).astype(np.float32)
        c = Co
    wts["identity_kernel"] = (rng.standard_normal((Ci, Co)) / np.sqrt(Ci)).astype(np.float32)
    wts["identity_bias"] = (0.1 * rng.standard_normal(Co)).astype(np.float32)
    leaves = {kk: torch.from_numpy(v).double().requires_grad_(True) for kk, v in wts.items()}
    xt = torch.from_numpy(x).double().requires_grad_(True)
    want = M.res_block(xt, leaves, k, acc_dtype=torch.float64)
    up = rng.standard_normal(tuple(want.shape)).astype(np.float32)
    want.backward(torch.from_numpy(up).double())
    blk = pkg.resBlock(Ci, Co)
    blk.build((B, h, w, Ci))
    blk.set_weights(wts)
    fv = pkg._flat.FlatVars(blk.owner_list(), "cuda")
    got = blk(torch.from_numpy(x).cuda(), save=True)
    assert rel_l2(got.cpu().numpy(), want.detach().numpy()) <= 2e-3
    dx = blk.backward(torch.from_numpy(up).cuda(), fv)
    assert rel_l2(dx.cpu().numpy(), xt.grad.numpy()) <= 2e-2
    assert rel_l2(fv.grad(blk.identity, "w").cpu().numpy().reshape(Ci, Co), leaves["identity_kernel"].grad.numpy()) <= 2e-3
    assert rel_l2(fv.grad(blk.conv1, "kernel").cpu().numpy(), leaves["conv1_kernel"].grad.numpy()) <= 2e-2
