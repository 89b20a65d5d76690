"""GPU parity of the sun pre-training step (train_sun.sun_train_step, train_sun.py:220-264) against autograd through the fp64 oracle.
Tolerances: loss 5e-3 relative; gradients relative L2 per variable — Dense kernels / biases 5e-2, conv kernels and norm parameters
1.5e-1 (TF32 forward and backward convs, ReLU masks and max-pool routing taken from TF32 activations: one flipped unit moves an early
layer's gradient by percent); the individual backward kernels are checked tightly in isolation first."""
import numpy as np
import pytest
import torch

from oracle import model_oracle as M

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_loss_backward_kernels(pkg):
    LIB, check = pkg._lib.LIB, pkg._lib.check
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(0)
    B, H, W = 2, 16, 32
    N = H * W
    p = torch.softmax(torch.from_numpy(rng.standard_normal((B, N)).astype(np.float32) * 2), -1)
    t = torch.softmax(torch.from_numpy(rng.standard_normal((B, N)).astype(np.float32) * 4), -1)
    pt = p.double().requires_grad_(True)
    loss = M.kl_divergence(t.double(), pt) + M.dog_l1(pt.reshape(B, H, W, 1), t.double().reshape(B, H, W, 1))
    loss.backward()
    pd, td = p.cuda(), t.cuda()
    g = torch.empty_like(pd)
    check(LIB.sky_kl_divergence_bwd(td.data_ptr(), pd.data_ptr(), g.data_ptr(), pd.numel(), 1.0 / B, 0, st))
    bp = torch.empty((B, 2 * H, 2 * W, 1), device="cuda")
    bg = torch.empty_like(bp)
    check(LIB.sky_dog_base(pd.data_ptr(), bp.data_ptr(), B, H, W, 1, st))
    check(LIB.sky_dog_base(td.data_ptr(), bg.data_ptr(), B, H, W, 1, st))
    db = torch.empty_like(bp)
    check(LIB.sky_dog_l1_bwd(bp.data_ptr(), bg.data_ptr(), db.data_ptr(), B, 2 * H, 2 * W, 1, 1.0 / bp.numel(), st))
    check(LIB.sky_dog_base_bwd(db.data_ptr(), g.data_ptr(), B, H, W, 1, 1, st))
    assert rel(g.cpu().numpy(), pt.grad.numpy()) < 1e-4, rel(g.cpu().numpy(), pt.grad.numpy())
    # softmax backward with the ReLU mask
    z = rng.standard_normal((B, N)).astype(np.float32)
    up = rng.standard_normal((B, N)).astype(np.float32)
    zt = torch.from_numpy(z).double().requires_grad_(True)
    a = torch.relu(zt)
    (torch.softmax(a, -1) * torch.from_numpy(up).double()).sum().backward()
    a32 = torch.relu(torch.from_numpy(z)).cuda()
    sm = pkg.sunpose_net.softmax(a32)
    gz = torch.empty_like(sm)
    upd = torch.from_numpy(up).cuda()
    check(LIB.sky_softmax_bwd_rows(sm.data_ptr(), upd.data_ptr(), a32.data_ptr(), gz.data_ptr(), B, N, st))
    assert rel(gz.cpu().numpy(), zt.grad.numpy()) < 1e-5
    # Dense weight gradient (ragged K and column tile, B > 32)
    for Bd, K, Nn in ((32, 100, 260), (40, 200, 512)):
        x = rng.standard_normal((Bd, K)).astype(np.float32)
        dy = rng.standard_normal((Bd, Nn)).astype(np.float32)
        dW = torch.empty((K, Nn), device="cuda")
        dbias = torch.empty(Nn, device="cuda")
        xd, dyd = torch.from_numpy(x).cuda(), torch.from_numpy(dy).cuda()      # keep the device copies alive across the launch
        check(LIB.sky_dense_bwd_filter(xd.data_ptr(), dyd.data_ptr(), dW.data_ptr(), dbias.data_ptr(), Bd, K, Nn, st))
        assert rel(dW.cpu().numpy(), x.astype(np.float64).T @ dy.astype(np.float64)) < 1e-5
        assert rel(dbias.cpu().numpy(), dy.astype(np.float64).sum(0)) < 1e-5
    # Adam: two steps against the Keras formula
    n = 1000
    w0, g1, g2 = (rng.standard_normal(n).astype(np.float32) for _ in range(3))
    w, m, v = torch.from_numpy(w0.copy()).cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    wr, mr, vr = w0.astype(np.float64), np.zeros(n), np.zeros(n)
    for step, gg in enumerate((g1, g2), start=1):
        ggd = torch.from_numpy(gg).cuda()
        check(LIB.sky_adam_step(w.data_ptr(), m.data_ptr(), v.data_ptr(), ggd.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-7, step, 1.0, st))
        mr = 0.9 * mr + 0.1 * gg
        vr = 0.999 * vr + 0.001 * gg.astype(np.float64) ** 2
        wr = wr - 1e-3 * np.sqrt(1 - 0.999 ** step) / (1 - 0.9 ** step) * mr / (np.sqrt(vr) + 1e-7)
    assert rel(w.cpu().numpy(), wr) < 1e-6


def test_smallc_weight_gradient(pkg):
    from oracle import da_oracle as O
    rng = np.random.default_rng(1)
    B, h, w, C, F, k = 2, 16, 64, 3, 32, 7
    x = rng.uniform(0, 1, (B, h, w, C)).astype(np.float32)
    kern = (rng.standard_normal((k * k * C, F)) / np.sqrt(k * k * C)).astype(np.float32)
    bias = rng.standard_normal(F).astype(np.float32)
    dy = rng.standard_normal((B, h, w, F)).astype(np.float32)
    _, want_dk, want_db = (g.numpy() for g in O.conv2d_backward(x, kern, bias, dy, k, acc_dtype=torch.float64))
    layer = pkg.conv2d(F, kernel_size=k, kernel_initializer=kern, bias_initializer=bias)
    xd = torch.from_numpy(x).cuda()
    layer(xd)
    dk, db = torch.empty_like(layer.kernel), torch.empty_like(layer.bias)
    dyd = torch.from_numpy(dy).cuda()
    pkg._lib.check(pkg._lib.LIB.sky_da_conv2d_smallc_bwd_filter(xd.data_ptr(), dyd.data_ptr(),
                                                                layer.offset_table.data_ptr(), dk.data_ptr(), db.data_ptr(), B, h, w, C, F, k,
                                                                torch.cuda.current_stream().cuda_stream))
    assert rel(dk.cpu().numpy(), want_dk) < 1e-4, rel(dk.cpu().numpy(), want_dk)
    assert rel(db.cpu().numpy(), want_db) < 1e-5


@pytest.mark.parametrize("mode,fwd_kernel,tol_fc,tol_conv", [("tf32", "strip", 8e-2, 1.5e-1), ("3xtf32", "band", 1e-3, 5e-3),
                                                             ("3xtf32", "strip", 1e-3, 5e-2)])
def test_sun_train_step_vs_autograd(pkg, monkeypatch, mode, fwd_kernel, tol_fc, tol_conv):
    """With `3xtf32` forward convs on the band-staged kernel the activations agree with the oracle to ~1e-6, no ReLU mask / arg-max flips
    occur, and what remains is the TF32 rounding of the backward convs (every backward kernel proven to <= 5e-3 end to end).  The
    row-strip forward kernel agrees to ~7e-6 per layer in `3xtf32` (more accumulation steps in TMEM): 2 of 65536 ReLU units of sunlayer3
    flip against the oracle, which moves the gradients below them by ~1.4e-2 (tools/dbg_strip_layer3.py) — a property of the
    discontinuity, the same backward kernels run.  With TF32 forward convs a handful of flipped units adds several percent (measured:
    fc2 1e-3, fc1 5e-2, convs 7-12e-2)."""
    monkeypatch.setattr(pkg.distortion_aware_ops, "DA_FORWARD_KERNEL", fwd_kernel)
    rng = np.random.default_rng(2)
    B, H, W = 2, 32, 128
    ldr = (np.round(255 * rng.uniform(0, 1, (B, H, W, 3))) / 255).astype(np.float32)
    gt = torch.softmax(torch.from_numpy(rng.standard_normal((B, H * W)).astype(np.float32) * 4), -1).numpy()
    ws = M.random_sunpose_weights(seed=5, H=H, W=W)
    net = pkg.sunpose_net.model(im_height=H, im_width=W, distortion_aware=True, math_mode=mode)
    tr = pkg.train_sun.SunTrainer(net, B, H, W, lr=1e-4)
    net.set_weights(ws)
    w_before = tr.flat_w.clone()
    pred, sungt, cams = tr.sun_train_step([None, torch.from_numpy(ldr).cuda()], torch.from_numpy(gt).cuda())
    want_loss, want = M.sun_train_step_grads(ldr, gt, ws)
    assert abs(float(tr.loss) / float(want_loss) - 1) < 5e-3, (float(tr.loss), float(want_loss))
    report = {}
    for name in ("fc1", "fc2"):
        layer = getattr(net, name)
        report[name + ".kernel"] = rel(tr._g(layer, "kernel").cpu().numpy(), want[name][0].numpy())
        report[name + ".bias"] = rel(tr._g(layer, "bias").cpu().numpy(), want[name][1].numpy())
    for name in ("sunlayer3", "sunlayer2", "sunlayer1"):
        layer = getattr(net, name)
        for i, (conv, norm) in enumerate(((layer.conv1, layer.norm1), (layer.conv2, layer.norm2)), start=1):
            report[f"{name}.conv{i}.kernel"] = rel(tr._g(conv, "kernel").cpu().numpy(), want[name][f"conv{i}_kernel"].numpy())
            report[f"{name}.conv{i}.bias"] = rel(tr._g(conv, "bias").cpu().numpy(), want[name][f"conv{i}_bias"].numpy())
            report[f"{name}.norm{i}.gamma"] = rel(tr._g(norm, "gamma").cpu().numpy(), want[name][f"norm{i}_gamma"].numpy())
            report[f"{name}.norm{i}.beta"] = rel(tr._g(norm, "beta").cpu().numpy(), want[name][f"norm{i}_beta"].numpy())
    import json, os
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({k: round(v, 5) for k, v in report.items()}, open(f"gpurun_out/sun_train_grad_report_{mode}_{fwd_kernel}.json", "w"), indent=1)
    for k, v in report.items():
        if k.endswith(".bias") and "conv" in k:
            continue      # a conv bias in front of an instance norm has an exactly-zero gradient: only noise on both sides
        assert v <= (tol_fc if k.startswith("fc") else tol_conv), (k, v)
    # Adam moved every weight by at most lr (first step: |m / sqrt(v)| <= 1), and did move them
    delta = (tr.flat_w - w_before).abs()
    assert float(delta.max()) <= 1.01e-4 and float(delta.mean()) > 1e-5
