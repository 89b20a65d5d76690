"""CPU checks of the adjoint formulas the train-step kernels implement (csrc/train_step_ops.cu, csrc/conv_bwd.cu), each restated in
numpy / torch exactly as the kernel comments state them and compared with autograd through the oracle.  They pin the MATH of the
backward kernels without a GPU; the kernels themselves are compared with the same oracle in tests/test_gpu_train_step.py."""
import numpy as np
import torch

from oracle import model_oracle as M


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_sun_radiance_and_maxnorm_adjoint():
    """sun_radiance_bwd_kernel + maxnorm_bwd_kernel + sunrad_heads_bwd_kernel vs autograd of sunrad_net.py:53-71 / generator.py:160,167."""
    rng = np.random.default_rng(0)
    B, hw, K, eps = 3, 40, 16, 1e-5
    sm = torch.softmax(torch.from_numpy(rng.standard_normal((B, hw)) * 3), -1).double().requires_grad_(True)
    flat = torch.from_numpy(rng.standard_normal((B, K))).double().requires_grad_(True)
    W = torch.from_numpy(rng.standard_normal((K, 2)) * 0.3).double().requires_grad_(True)
    bias = torch.from_numpy(rng.standard_normal(2) * 0.1).double().requires_grad_(True)
    up = torch.from_numpy(rng.standard_normal((B, hw, 3))).double()
    sqrt_pi = float(np.sqrt(np.float32(np.pi)))
    gb = flat @ W + bias
    gam, bet = torch.sigmoid(gb[:, 0:1]), torch.sigmoid(gb[:, 1:2])
    x = sm / sm.max()
    v = torch.exp(-((1 - x) ** 2) / (bet + eps)) * gam / (bet * sqrt_pi + eps)
    v = v * 2000.0                                        # push part of the map over the clamp
    vc = torch.where(v > 30000.0, torch.full_like(v, 30000.0), v)
    out3 = (torch.log(1 + 10 * vc) / np.log(11.0)).unsqueeze(-1).expand(B, hw, 3)
    (out3 * up).sum().backward()
    # --- kernel restatement (with the same x2000 factor folded into gamma's scale) ---
    smn, gbn = sm.detach().numpy(), gb.detach().numpy()
    mx = smn.max()
    g_, b_ = 1 / (1 + np.exp(-gbn[:, 0:1])), 1 / (1 + np.exp(-gbn[:, 1:2]))
    den1, den2 = b_ + eps, b_ * sqrt_pi + eps
    xx = smn / mx
    d = 1 - xx
    ex = np.exp(-(d * d) / den1)
    vv = ex * g_ / den2 * 2000.0
    vcl = np.minimum(vv, 30000.0)
    dsum = up.numpy().sum(-1)
    dv = dsum * 10.0 / ((1 + 10 * vcl) * np.log(11.0))
    dv = np.where(vv > 30000.0, 0.0, dv)
    du = dv * vv
    dx = du * 2 * d / den1
    a_g = (dv * ex / den2 * 2000.0).sum(1, keepdims=True)
    a_b = (du * d * d / (den1 * den1) - dv * vv * sqrt_pi / den2).sum(1, keepdims=True)
    dgb = np.concatenate([a_g * g_ * (1 - g_), a_b * b_ * (1 - b_)], 1)
    dot, cnt = (dx * xx).sum(), (smn == mx).sum()
    dsm = dx / mx - (smn == mx) * dot / (mx * cnt)
    assert rel(dsm, sm.grad.numpy()) < 1e-10
    assert rel(flat.detach().numpy().T @ dgb, W.grad.numpy()) < 1e-10
    assert rel(dgb.sum(0), bias.grad.numpy()) < 1e-10
    assert rel(dgb @ W.detach().numpy().T, flat.grad.numpy()) < 1e-10


def test_train_tail_adjoint():
    """train_tail_fwd_kernel / train_tail_bwd_kernel vs autograd of generator.py:120-124,150-156 + train.py:289-298,324."""
    rng = np.random.default_rng(1)
    n, thr, slope = 200, 0.12, 0.1
    c_sky = torch.from_numpy(rng.standard_normal((n, 3)) * 0.5).double().requires_grad_(True)
    c_sun = torch.from_numpy(rng.standard_normal((n, 3)) * 0.5).double().requires_grad_(True)
    srg = torch.from_numpy(rng.uniform(0, 1.5, (n, 3))).double().requires_grad_(True)
    ldr = torch.from_numpy(rng.uniform(0, 1, (n, 3))).double()
    hdr_t = torch.from_numpy(rng.uniform(0, 3, (n, 3))).double()
    g_dog, g_dis8, g_vgg4 = (torch.from_numpy(rng.standard_normal((n, c))).double() for c in (3, 8, 4))
    w_l1, w_vgg = 0.37, 255.0
    s = torch.relu(ldr + M.leaky_relu(c_sky, slope))
    u = torch.relu(srg + M.leaky_relu(c_sun, slope))
    with torch.no_grad():
        alpha = torch.clamp((M.hdr_log_decompression(s).amax(1, keepdim=True) - 1 + thr).clamp(min=0) / thr, max=1)
    y_gamma = (1 - alpha) * s + alpha * u
    y_lin = M.hdr_log_decompression(y_gamma)
    loss = (y_lin * (g_dog + g_dis8[:, 3:6])).sum() + w_l1 * (y_lin - hdr_t).abs().sum() + w_vgg * (y_gamma * g_vgg4[:, :3]).sum()
    loss.backward()
    assert 0.05 < float((alpha > 0).double().mean()) and float((alpha < 1).double().mean()) > 0.05      # the ramp is exercised
    yl, a = y_lin.detach().numpy(), alpha.numpy()
    gl = w_l1 * np.sign(yl - hdr_t.numpy()) + g_dog.numpy() + g_dis8.numpy()[:, 3:6]
    gy = gl * (np.log(11.0) * (10 * yl + 1) / 10) + w_vgg * g_vgg4.numpy()[:, :3]
    cs, cu = c_sky.detach().numpy(), c_sun.detach().numpy()
    sv = ldr.numpy() + np.where(cs > 0, cs, slope * cs)
    uv = srg.detach().numpy() + np.where(cu > 0, cu, slope * cu)
    gs = np.where(sv > 0, (1 - a) * gy, 0.0)
    gu = np.where(uv > 0, a * gy, 0.0)
    assert rel(gs * np.where(cs > 0, 1, slope), c_sky.grad.numpy()) < 1e-12
    assert rel(gu * np.where(cu > 0, 1, slope), c_sun.grad.numpy()) < 1e-12
    assert rel(gu, srg.grad.numpy()) < 1e-12


def test_batch_norm_train_adjoint_with_groups():
    """bn_bwd_reduce_kernel / bn_bwd_apply_kernel vs autograd of two independent Keras BatchNormalization calls (real / generated halves)."""
    rng = np.random.default_rng(2)
    B, hw, F, groups, eps, slope = 4, 6, 5, 2, 1e-3, 0.3
    x = torch.from_numpy(rng.standard_normal((B, hw, F)) * 2 + 1).double().requires_grad_(True)
    gamma = torch.from_numpy(1 + 0.2 * rng.standard_normal(F)).double().requires_grad_(True)
    beta = torch.from_numpy(0.2 * rng.standard_normal(F)).double().requires_grad_(True)
    up = torch.from_numpy(rng.standard_normal((B, hw, F))).double()
    halves = []
    for g in range(groups):
        xs = x[g * B // groups:(g + 1) * B // groups].reshape(1, -1, 1, F)
        y, _, _ = M.batch_norm_train(xs.reshape(B // groups, hw, 1, F), gamma, beta, eps)
        halves.append(M.leaky_relu(y.reshape(B // groups, hw, F), slope))
    act = torch.cat(halves, 0)
    (act * up).sum().backward()
    xn, an = x.detach().numpy().reshape(groups, -1, F), act.detach().numpy().reshape(groups, -1, F)
    dy = up.numpy().reshape(groups, -1, F) * np.where(an > 0, 1, slope)
    mean, var = xn.mean(1, keepdims=True), xn.var(1, keepdims=True)
    rstd = 1 / np.sqrt(var + eps)
    xh = (xn - mean) * rstd
    m1, m2 = dy.mean(1, keepdims=True), (dy * xh).mean(1, keepdims=True)
    dx = gamma.detach().numpy() * rstd * (dy - m1 - xh * m2)
    assert rel(dx.reshape(B, hw, F), x.grad.numpy()) < 1e-10
    assert rel((dy * xh).sum((0, 1)), gamma.grad.numpy()) < 1e-10
    assert rel(dy.sum((0, 1)), beta.grad.numpy()) < 1e-10


def _transposed_conv_via_forward(dy, w4, h, w, stride):
    """sky_conv2d_bwd_data's formulation: a forward pass over dy with the flipped, transposed kernel and the transposed sampler of
    csrc/da_conv_fwd.cu (sample_corners, plain == 2): tap (a, b) of gradient pixel (i, j) reads dy[(i + a - P) / s, (j + b - Q) / s]
    when both divisions are exact, P = k - 1 - (SAME front pad)."""
    B, oh, ow, F = dy.shape
    k, _, C, _ = w4.shape

    def front(n, o):
        return max((o - 1) * stride + k - n, 0) // 2

    P, Q = k - 1 - front(h, oh), k - 1 - front(w, ow)
    wt = w4[::-1, ::-1].transpose(0, 1, 3, 2)              # [k, k, F, C], taps flipped
    dx = np.zeros((B, h, w, C))
    for a in range(k):
        for b in range(k):
            for i in range(h):
                ty = i + a - P
                if ty < 0 or ty % stride or ty // stride >= oh:
                    continue
                for j in range(w):
                    tx = j + b - Q
                    if tx < 0 or tx % stride or tx // stride >= ow:
                        continue
                    dx[:, i, j, :] += dy[:, ty // stride, tx // stride, :] @ wt[a, b]
    return dx


def test_data_gradient_as_forward_pass():
    rng = np.random.default_rng(3)
    for (h, w, C, F, k, s) in ((6, 8, 3, 4, 3, 1), (6, 8, 3, 4, 3, 2), (8, 8, 2, 3, 4, 2), (5, 7, 2, 3, 4, 1), (7, 9, 2, 2, 7, 1), (7, 5, 2, 2, 3, 2)):
        x = torch.from_numpy(rng.standard_normal((2, h, w, C))).double().requires_grad_(True)
        w4 = rng.standard_normal((k, k, C, F))
        y = M.conv2d_same(x, torch.from_numpy(w4).double(), torch.zeros(F).double(), stride=s, acc_dtype=torch.float64)
        dy = rng.standard_normal(tuple(y.shape))
        y.backward(torch.from_numpy(dy))
        assert rel(_transposed_conv_via_forward(dy, w4, h, w, s), x.grad.numpy()) < 1e-12, (h, w, C, F, k, s)


def test_resize_adjoint_weights():
    """resize_bilinear_bwd_kernel scatters with (1-ly)(1-lx), (1-ly)lx, ly(1-lx), ly lx to (ylo|yhi, xlo|xhi): the transpose of the oracle's resize."""
    from oracle import da_oracle as O
    rng = np.random.default_rng(4)
    for (h, w, oh, ow) in ((4, 6, 8, 12), (3, 5, 6, 10), (8, 32, 16, 64)):
        x = torch.from_numpy(rng.standard_normal((1, h, w, 2))).double().requires_grad_(True)
        y = O.resize_bilinear(x, oh, ow)
        dy = rng.standard_normal(tuple(y.shape))
        y.backward(torch.from_numpy(dy))
        dx = np.zeros((1, h, w, 2))
        sy, sx = np.float32(h) / np.float32(oh), np.float32(w) / np.float32(ow)
        for oy in range(oh):
            fy = (np.float32(oy) + np.float32(0.5)) * sy - np.float32(0.5)
            ylo, yhi, ly = max(int(np.floor(fy)), 0), min(int(np.ceil(fy)), h - 1), float(fy - np.floor(fy))
            for ox in range(ow):
                fx = (np.float32(ox) + np.float32(0.5)) * sx - np.float32(0.5)
                xlo, xhi, lx = max(int(np.floor(fx)), 0), min(int(np.ceil(fx)), w - 1), float(fx - np.floor(fx))
                for yy, wy in ((ylo, 1 - ly), (yhi, ly)):
                    for xx_, wx in ((xlo, 1 - lx), (xhi, lx)):
                        dx[0, yy, xx_] += wy * wx * dy[0, oy, ox]
        assert rel(dx, x.grad.numpy()) < 1e-6


def test_tf32_emulation_rounding():
    from oracle import tf32_emu as E
    v = torch.tensor([1.0, 1.0 + 2 ** -11, 1.0 + 2 ** -11 + 2 ** -20, -1.0 - 2 ** -11, 3.14159265, 1e-30, 0.0], dtype=torch.float32)
    r = E.round_tf32(v)
    assert r[0] == 1.0 and r[1] == 1.0 + 2 ** -10 and r[2] == 1.0 + 2 ** -10 and r[3] == -1.0 - 2 ** -10       # ties away from zero
    assert abs(float(r[4]) - 3.14159265) < 3.14159265 * 2 ** -11 and r[6] == 0.0
    assert (r.view(torch.int32) & 0x1FFF).abs().sum() == 0
    with E.emulate():
        a = torch.randn(8, 16, dtype=torch.float64, requires_grad=True)
        b = torch.randn(16, 4, dtype=torch.float64, requires_grad=True)
        y = E.matmul(a, b)
        y.sum().backward()
        assert torch.allclose(y, E.round_tf32(a) @ E.round_tf32(b))
        assert a.grad is not None and b.grad is not None
