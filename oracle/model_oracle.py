"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/da_oracle.py for the rules).

CPU (torch fp32 / fp64) restatement of the residual trunk of the reference generator with the distortion-aware convs
wired in as the commented lines generator.py:14,18 do:

  instance_norm()   tfa.layers.InstanceNormalization defaults (generator.py:15,19): per (n, c) over H x W, biased variance
                    (tf.nn.moments), eps 1e-3, y = x * (gamma * rsqrt(var + eps)) + (beta - mean * gamma * rsqrt(var + eps))
                    — TensorFlow-Addons is not vendored in the reference; restated from its published GroupNormalization code
  res_block()       generator.py:26-35
  res_layer()       generator.py:46-49
and of the rest of the path (sun branch, train / test tail, sun pre-train step) further down.

Pinning: the distortion-aware sampling against tests/golden/da_golden.npz (the reference's own distortion_aware_ops.py executed over a
numpy TensorFlow shim); hdr_log_*, apply_rf, ldr_synth and the sunRadNet radiance tail against tests/golden/utils_golden.npz (the
reference's own tf_utils.py / sunrad_net.py / train.py over the same shim); the Keras / TFA semantics that are not in /root/reference (instance norm, batch norm,
SAME padding, max-pool routing, KLDivergence, gaussian_filter2d) against numpy / scipy / closed forms in tests/test_oracle_model.py.
With respect to a real TensorFlow run parity remains unpinned (TensorFlow cannot be installed in the build container).
"""
from __future__ import annotations

import numpy as np
import torch

from . import da_oracle as O
from . import tf32_emu as emu


def instance_norm(x, gamma, beta, eps=1e-3):
    mean = x.mean(dim=(1, 2), keepdim=True)
    var = ((x - mean) ** 2).mean(dim=(1, 2), keepdim=True)
    inv = torch.rsqrt(var + eps) * gamma
    return x * inv + (beta - mean * inv)


def leaky_relu(x, alpha):
    return torch.where(x > 0, x, x * alpha)


def res_block(x, w, k=3, dilation=1, acc_dtype=torch.float32):
    """w: dict with conv1_kernel, conv1_bias, norm1_gamma, norm1_beta, conv2_*, norm2_* (torch tensors)."""
    dt = acc_dtype
    c1 = O.conv2d_forward(x, w["conv1_kernel"], w["conv1_bias"], k, dilation, acc_dtype=dt)
    a1 = leaky_relu(instance_norm(c1, w["norm1_gamma"].to(dt), w["norm1_beta"].to(dt)), 0.1)
    c2 = O.conv2d_forward(a1 if dt == torch.float32 else a1, w["conv2_kernel"], w["conv2_bias"], k, dilation, acc_dtype=dt)
    n2 = instance_norm(c2, w["norm2_gamma"].to(dt), w["norm2_beta"].to(dt))
    ident = x.to(dt)
    if "identity_kernel" in w:                 # generator.py:23-24: 1x1 projection shortcut when the channel count changes
        ik = O._as_t(w["identity_kernel"])
        ident = conv2d_same(x, ik.reshape(1, 1, x.shape[-1], -1), w["identity_bias"], acc_dtype=dt)
    return ident + n2


def res_layer(x, blocks, k=3, dilation=1, acc_dtype=torch.float32):
    x = O._as_t(x).to(acc_dtype)
    for w in blocks:
        x = res_block(x, w, k, dilation, acc_dtype)
    return x


def conv2d_same(x, w4, b, stride=1, acc_dtype=torch.float32, tf32=None):
    """tf.nn.conv2d(x, w, strides, 'SAME') + bias_add (ops.py:41-42) on NHWC input, HWIO kernel.  TensorFlow SAME padding:
    out = ceil(n/s), total = max((out-1)*s + k - n, 0), floor(total/2) in front, the rest behind."""
    x, w4, b = O._as_t(x).to(acc_dtype), O._as_t(w4).to(acc_dtype), O._as_t(b).to(acc_dtype)
    k = w4.shape[0]
    B, h, w, C = x.shape
    pads = []
    for n in (h, w):
        out = -(-n // stride)
        total = max((out - 1) * stride + k - n, 0)
        pads.append((total // 2, total - total // 2))
    xp = torch.nn.functional.pad(x.permute(0, 3, 1, 2), (pads[1][0], pads[1][1], pads[0][0], pads[0][1]))
    if tf32 is None:      # layers the library runs on the fp32 pipes: image-like inputs (conv1_d) and 3-filter outputs (conv1_f / conv1_u)
        F = w4.shape[3]
        tf32 = not (C <= 4 and F <= 32 and stride == 1 and k % 2 == 1)
        if F <= 4 and stride == 1 and k % 2 == 1:
            tf32 = "weights"
    y = emu.conv2d(xp, w4.permute(3, 2, 0, 1), stride=stride, tf32=tf32) + b.reshape(1, -1, 1, 1)
    return y.permute(0, 2, 3, 1).contiguous()


def hdr_log_decompression(x, valid_dr=10.0):
    """tf_utils.hdr_logDecompression (tf_utils.py:273-280)."""
    den = torch.log(torch.tensor(1.0 + valid_dr, dtype=x.dtype))
    return (torch.exp(x * den) - 1.0) / valid_dr


def sky_branch(ldr, w, k=3, acc_dtype=torch.float32, decompress=True):
    """inference.py:84-86 with the distortion-aware trunk: encode (generator.py:92-108) -> sky_decode (:110-125) ->
    hdr_logDecompression.  w: the dict layout of <package>.generator.model.set_weights."""
    dt = acc_dtype
    x = O._as_t(ldr).to(dt)
    inp = x

    def norm_act(y, name):
        g, b = (O._as_t(v).to(dt) for v in w[name])
        return leaky_relu(instance_norm(y, g, b), 0.1)

    x = norm_act(conv2d_same(x, *w["conv1_d"], stride=1, acc_dtype=dt), "norm1_d")
    x = norm_act(conv2d_same(x, *w["conv2_d"], stride=2, acc_dtype=dt), "norm2_d")
    x = norm_act(conv2d_same(x, *w["conv3_d"], stride=2, acc_dtype=dt), "norm3_d")
    blocks = [{kk: O._as_t(v) for kk, v in blk.items()} for blk in w["res"]]
    x = res_layer(x, blocks, k, acc_dtype=dt)
    H, W = inp.shape[1], inp.shape[2]
    x = O.resize_bilinear(x, H // 2, W // 2)                                   # ops.py:122
    x = norm_act(conv2d_same(x, *w["conv3_f"], acc_dtype=dt), "norm3_f")
    x = O.resize_bilinear(x, H, W)
    x = norm_act(conv2d_same(x, *w["conv2_f"], acc_dtype=dt), "norm2_f")
    sky = leaky_relu(conv2d_same(x, *w["conv1_f"], acc_dtype=dt), 0.1)        # generator.py:120-121
    sky = torch.relu(inp + sky)                                                # :123-124
    return hdr_log_decompression(sky) if decompress else sky


def maxpool2x2_same(x):
    """tf.nn.max_pool(ksize 2, strides 2, 'SAME') on NHWC (ops.py:299-300): out = ceil(n/2), -inf padding at the far edge."""
    B, h, w, C = x.shape
    xp = torch.nn.functional.pad(x.permute(0, 3, 1, 2), (0, w % 2, 0, h % 2), value=float("-inf"))
    return torch.nn.functional.max_pool2d(xp, 2, 2).permute(0, 2, 3, 1).contiguous()


def sunpose_layer(x, w, k, distortion_aware=True, acc_dtype=torch.float32):
    """sunposeLayer.call (sunpose_net.py:19-30): (conv -> IN -> relu) x 2; conv = the distortion-aware layer (:11,16) or the
    plain SAME conv (:10,15).  w: conv{1,2}_kernel [k*k*C, F], conv{1,2}_bias, norm{1,2}_gamma / _beta."""
    dt = acc_dtype
    for i in (1, 2):
        kern, bias = O._as_t(w[f"conv{i}_kernel"]), O._as_t(w[f"conv{i}_bias"])
        if distortion_aware:
            x = O.conv2d_forward(x, kern, bias, k, acc_dtype=dt)
        else:
            C = kern.shape[0] // (k * k)
            x = conv2d_same(x, kern.reshape(k, k, C, -1), bias, acc_dtype=dt)
        x = torch.relu(instance_norm(x, O._as_t(w[f"norm{i}_gamma"]).to(dt), O._as_t(w[f"norm{i}_beta"]).to(dt)))
    return x


def sunpose_estimation(x, w, distortion_aware=True, acc_dtype=torch.float32):
    """sunpose_net.model.sunposeEstimation (sunpose_net.py:54-72) -> (softmax [B, H*W], [act1, act2, act3])."""
    dt = acc_dtype
    x = O._as_t(x).to(dt)
    acts = []
    for name, k in (("sunlayer1", 7), ("sunlayer2", 3), ("sunlayer3", 3)):
        a = sunpose_layer(x, w[name], k, distortion_aware, dt)
        acts.append(a)
        x = maxpool2x2_same(a)
    flat = x.reshape(x.shape[0], -1)                                    # Keras Flatten: (h, w, c)
    h1 = torch.relu(flat @ O._as_t(w["fc1"][0]).to(dt) + O._as_t(w["fc1"][1]).to(dt))
    h2 = torch.relu(h1 @ O._as_t(w["fc2"][0]).to(dt) + O._as_t(w["fc2"][1]).to(dt))
    return torch.softmax(h2, dim=-1), acts


def random_sunpose_weights(seed, H, W, affine_noise=True):
    import numpy as np
    rng = np.random.default_rng(seed)
    w = {}
    cin = 3
    for name, f, k in (("sunlayer1", 32, 7), ("sunlayer2", 64, 3), ("sunlayer3", 128, 3)):
        d = {}
        c = cin
        for i in (1, 2):
            lim = (6.0 / (k * k * c + f)) ** 0.5
            d[f"conv{i}_kernel"] = rng.uniform(-lim, lim, (k * k * c, f)).astype(np.float32)
            d[f"conv{i}_bias"] = (0.05 * rng.standard_normal(f) * affine_noise).astype(np.float32)
            d[f"norm{i}_gamma"] = (1 + 0.1 * rng.standard_normal(f) * affine_noise).astype(np.float32)
            d[f"norm{i}_beta"] = (0.1 * rng.standard_normal(f) * affine_noise).astype(np.float32)
            c = f
        w[name] = d
        cin = f
    flat = (H // 8) * (W // 8) * 128
    fc = H * W
    for name, kin in (("fc1", flat), ("fc2", fc)):
        lim = (6.0 / (kin + fc)) ** 0.5
        w[name] = (rng.uniform(-lim, lim, (kin, fc)).astype(np.float32), (0.05 * rng.standard_normal(fc) * affine_noise).astype(np.float32))
    return w


def random_generator_weights(seed, k=3, affine_noise=True):
    """numpy-seeded weights for the sky branch with the reference's initialiser distributions (Keras glorot_uniform on
    the 4-D kernels), biases / gamma / beta optionally perturbed so they are exercised."""
    import numpy as np
    rng = np.random.default_rng(seed)
    w = {}

    def conv(name, kk, cin, cout):
        lim = (6.0 / (kk * kk * cin + kk * kk * cout)) ** 0.5
        w[name] = (rng.uniform(-lim, lim, (kk, kk, cin, cout)).astype(np.float32),
                   (0.05 * rng.standard_normal(cout) * affine_noise).astype(np.float32))

    def norm(name, c):
        w[name] = ((1 + 0.1 * rng.standard_normal(c) * affine_noise).astype(np.float32),
                   (0.1 * rng.standard_normal(c) * affine_noise).astype(np.float32))

    conv("conv1_d", 7, 3, 32); norm("norm1_d", 32)
    conv("conv2_d", 3, 32, 64); norm("norm2_d", 64)
    conv("conv3_d", 3, 64, 128); norm("norm3_d", 128)
    w["res"] = [{kk: v.numpy() for kk, v in blk.items()} for blk in random_trunk_weights(6, 128, k, seed + 1, affine_noise)]
    conv("conv3_f", 3, 128, 64); norm("norm3_f", 64)
    conv("conv2_f", 3, 64, 32); norm("norm2_f", 32)
    conv("conv1_f", 7, 32, 3)
    return w


def random_trunk_weights(n_blocks, C, k, seed, affine_noise=True):
    """numpy-seeded weights with the reference's initialiser distributions (glorot_uniform kernels, zero bias, gamma=1,
    beta=0), optionally perturbed so bias / gamma / beta are exercised."""
    import numpy as np
    rng = np.random.default_rng(seed)
    blocks = []
    lim = (6.0 / (k * k * C + C)) ** 0.5
    for _ in range(n_blocks):
        w = {}
        for i in (1, 2):
            w[f"conv{i}_kernel"] = torch.from_numpy(rng.uniform(-lim, lim, (k * k * C, C)).astype(np.float32))
            w[f"conv{i}_bias"] = torch.from_numpy((0.1 * rng.standard_normal(C) * affine_noise).astype(np.float32))
            w[f"norm{i}_gamma"] = torch.from_numpy((1 + 0.1 * rng.standard_normal(C) * affine_noise).astype(np.float32))
            w[f"norm{i}_beta"] = torch.from_numpy((0.1 * rng.standard_normal(C) * affine_noise).astype(np.float32))
        blocks.append(w)
    return blocks


# ---------------------------------------------------------------------------------------------------------------------
# sun branch of generator inference (inference.py:81-112)
# ---------------------------------------------------------------------------------------------------------------------
def hdr_log_compression(x, valid_dr=10.0):
    """tf_utils.hdr_logCompression (tf_utils.py:263-271)."""
    return torch.log(1.0 + valid_dr * x) / torch.log(torch.tensor(1.0 + valid_dr, dtype=x.dtype))


def grad_cam_layer(y_c, A_k):
    """grad_cam.layer (grad_cam.py:29-45): tf.gradients(y_c, A_k) sums y_c over the batch; the samples are independent."""
    grad = torch.autograd.grad(y_c.sum(), A_k, retain_graph=True)[0]
    weights = grad.mean(dim=(1, 2))                                       # :34
    cam = torch.einsum('bc,bwhc->bwh', weights, A_k)                      # :35
    return torch.relu(cam).unsqueeze(-1)                                  # :38-43


def batch_norm_inference(x, gamma, beta, mean, var, eps=1e-3):
    """Keras BatchNormalization with training=False: moving statistics, epsilon 1e-3."""
    inv = gamma * torch.rsqrt(var + eps)
    return x * inv + (beta - mean * inv)


def batch_norm_train(x, gamma, beta, eps=1e-3):
    """Keras BatchNormalization with training=True (sunrad_net.py:25, discriminator.py:24): batch statistics over (N, H, W), biased
    variance to normalise.  Returns (y, batch mean, biased batch variance)."""
    mean = x.mean(dim=(0, 1, 2))
    var = ((x - mean) ** 2).mean(dim=(0, 1, 2))
    inv = gamma * torch.rsqrt(var + eps)
    return x * inv + (beta - mean * inv), mean.detach(), var.detach()


def bn_moving_update(moving_mean, moving_var, mean, var_biased, count, momentum=0.99):
    """The moving-statistics assignment of a training-mode call (Keras BatchNormalization on the fused path, momentum 0.99):
    TensorFlow's FusedBatchNormV3 hands back the Bessel-corrected batch variance for the running average (restated from the
    published kernel, not in /root/reference)."""
    unbiased = var_biased * (count / max(count - 1, 1))
    return moving_mean * momentum + mean * (1 - momentum), moving_var * momentum + unbiased * (1 - momentum)


def down_blocks(h, w, dt, training=False, bn_stats=None):
    """d1..d4 of sunrad_net.sunRadNet / discriminator.model: Conv2D(4x4, SAME, no bias) -> BatchNormalization -> LeakyReLU(0.3).
    training=True: batch statistics; bn_stats (dict) then receives name -> (mean, biased variance, element count per channel)."""
    for name, stride in (("d1", 2), ("d2", 2), ("d3", 2), ("d4", 1)):
        d = w[name]
        kern = O._as_t(d["kernel"]).to(dt)
        h = conv2d_same(h, kern, torch.zeros(kern.shape[-1], dtype=dt), stride=stride, acc_dtype=dt)     # use_bias=False
        if "gamma" in d:
            if training:
                h, mean, var = batch_norm_train(h, O._as_t(d["gamma"]).to(dt), O._as_t(d["beta"]).to(dt))
                if bn_stats is not None:
                    bn_stats[name] = (mean, var, h.shape[0] * h.shape[1] * h.shape[2])
            else:
                h = batch_norm_inference(h, *(O._as_t(d[k]).to(dt) for k in ("gamma", "beta", "moving_mean", "moving_variance")))
        h = leaky_relu(h, 0.3)                                              # Keras LeakyReLU() default alpha
    return h


def sunrad_net(x, actv_map, w, eps=1e-5, acc_dtype=torch.float32, return_heads=False, training=False, bn_stats=None):
    """sunRadNet.call (sunrad_net.py:46-71).  w: d1..d4 -> dict(kernel [4,4,C,F] (, gamma, beta, moving_mean,
    moving_variance)), gamma / beta -> (kernel [flat, 1], bias [1])."""
    dt = acc_dtype
    h = down_blocks(actv_map.to(dt), w, dt, training, bn_stats)
    flat = h.reshape(h.shape[0], -1)
    gamma = flat @ O._as_t(w["gamma"][0]).to(dt) + O._as_t(w["gamma"][1]).to(dt)
    beta = flat @ O._as_t(w["beta"][0]).to(dt) + O._as_t(w["beta"][1]).to(dt)
    gamma_in = torch.sigmoid(gamma).reshape(-1, 1, 1, 1)
    beta_in = torch.sigmoid(beta).reshape(-1, 1, 1, 1)
    v = -((1.0 - x.to(dt)) ** 2.0)
    v = v / (beta_in + eps)
    v = torch.exp(v)
    v = v * gamma_in
    const = beta_in * torch.sqrt(torch.tensor(float(np.float32(np.pi)), dtype=dt))
    v = v / (const + eps)
    v = torch.where(v > 30000.0, torch.full_like(v, 30000.0), v)
    return (v, gamma_in, beta_in) if return_heads else v


def decode_branch(x, inp_add, w, names, dt):
    """sky_decode / sun_decode (generator.py:110-156): two resize-deconv + IN + lrelu, 7x7 conv, lrelu, + addend, relu."""
    c3, n3, c2, n2, c1 = names
    H, W = inp_add.shape[1], inp_add.shape[2]

    def norm_act(y, name):
        g, b = (O._as_t(v).to(dt) for v in w[name])
        return leaky_relu(instance_norm(y, g, b), 0.1)

    x = O.resize_bilinear(x, H // 2, W // 2)
    x = norm_act(conv2d_same(x, *w[c3], acc_dtype=dt), n3)
    x = O.resize_bilinear(x, H, W)
    x = norm_act(conv2d_same(x, *w[c2], acc_dtype=dt), n2)
    y = leaky_relu(conv2d_same(x, *w[c1], acc_dtype=dt), 0.1)
    return torch.relu(inp_add + y)


def generator_inference(ldr, wg, ws, k=3, threshold=0.12, distortion_aware_sunpose=True, acc_dtype=torch.float32, details=False):
    """inference.generator_in_step (inference.py:81-112).  wg: generator weights (random_generator_weights(..., sun=True)),
    ws: sun-position weights (random_sunpose_weights)."""
    import numpy as np
    dt = acc_dtype
    inp = O._as_t(ldr).to(dt)
    B, H, W, _ = inp.shape

    def norm_act(y, name):
        g, b = (O._as_t(v).to(dt) for v in wg[name])
        return leaky_relu(instance_norm(y, g, b), 0.1)

    x = norm_act(conv2d_same(inp, *wg["conv1_d"], stride=1, acc_dtype=dt), "norm1_d")          # encode, :83
    x = norm_act(conv2d_same(x, *wg["conv2_d"], stride=2, acc_dtype=dt), "norm2_d")
    x = norm_act(conv2d_same(x, *wg["conv3_d"], stride=2, acc_dtype=dt), "norm3_d")
    blocks = [{kk: O._as_t(v) for kk, v in blk.items()} for blk in wg["res"]]
    res_out = res_layer(x, blocks, k, acc_dtype=dt)
    sky_gamma = decode_branch(res_out, inp, wg, ("conv3_f", "norm3_f", "conv2_f", "norm2_f", "conv1_f"), dt)   # :84
    sky_lin = hdr_log_decompression(sky_gamma)                                                 # :85

    x_sun = inp.clone().requires_grad_(True)           # so that the activation maps are part of an autograd graph
    sm, acts = sunpose_estimation(x_sun, ws, distortion_aware=distortion_aware_sunpose, acc_dtype=dt)   # :87
    sunpose_pred = sm.reshape(B, H, W, 1)
    alpha = sky_lin.amax(dim=3)                                                                # :91
    alpha = torch.clamp((alpha - 1.0 + threshold).clamp(min=0.0) / threshold, max=1.0)         # :92
    alpha_c3 = alpha.unsqueeze(-1).expand(B, H, W, 3)
    y_c = sm.amax(dim=1)                                                                       # :98 (ties share the gradient, like TF)
    cams = [grad_cam_layer(y_c, a).detach() for a in acts]                                     # :100-102
    sm, sunpose_pred = sm.detach(), sunpose_pred.detach()
    normed = sunpose_pred / sunpose_pred.max()                                                 # generator.py:160
    r2 = O.resize_bilinear(cams[1], H, W)                                                      # :161-162
    r3 = O.resize_bilinear(cams[2], H, W)
    plz = torch.cat([inp, cams[0], r2, r3], dim=-1)                                            # :164
    sun_rad = sunrad_net(normed, plz, wg["sun"], acc_dtype=dt)                                 # :165
    sun_rad_lin = sun_rad.expand(B, H, W, 3)                                                   # :167
    sun_rad_gamma = hdr_log_compression(sun_rad_lin)                                           # inference.py:105
    sun_gamma = decode_branch(res_out, sun_rad_gamma, wg, ("conv3_u", "norm3_u", "conv2_u", "norm2_u", "conv1_u"), dt)   # :106
    y_gamma = (1.0 - alpha_c3) * sky_gamma + alpha_c3 * sun_gamma                              # :108-110
    y_lin = hdr_log_decompression(y_gamma)                                                     # :111
    if details:
        return dict(y_lin=y_lin, sky_gamma=sky_gamma, sm=sm, cams=cams, sun_rad_gamma=sun_rad_gamma, sun_gamma=sun_gamma,
                    alpha=alpha, acts=[a.detach() for a in acts], plz=plz)
    return y_lin


def random_sunrad_weights(seed, H, W, bn_noise=True):
    import numpy as np
    rng = np.random.default_rng(seed)
    w = {}
    cin = 6
    for name, f, norm in (("d1", 64, False), ("d2", 128, True), ("d3", 256, True), ("d4", 512, True)):
        d = {"kernel": (0.02 * rng.standard_normal((4, 4, cin, f))).astype(np.float32)}
        if norm:
            d["gamma"] = (1 + 0.1 * rng.standard_normal(f) * bn_noise).astype(np.float32)
            d["beta"] = (0.1 * rng.standard_normal(f) * bn_noise).astype(np.float32)
            d["moving_mean"] = (0.01 * rng.standard_normal(f) * bn_noise).astype(np.float32)
            d["moving_variance"] = (1 + 0.2 * rng.uniform(-1, 1, f) * bn_noise).astype(np.float32)
        w[name] = d
        cin = f
    flat = (H // 8) * (W // 8) * 512
    lim = (6.0 / (flat + 1)) ** 0.5
    for head in ("gamma", "beta"):
        w[head] = (rng.uniform(-lim, lim, (flat, 1)).astype(np.float32), (0.1 * rng.standard_normal(1) * bn_noise).astype(np.float32))
    return w


def random_full_generator_weights(seed, H, W, k=3, affine_noise=True):
    """random_generator_weights plus the sun decoder (conv3_u, conv2_u, conv1_u, norms) and sunRadNet."""
    import numpy as np
    w = random_generator_weights(seed, k, affine_noise)
    rng = np.random.default_rng(seed + 77)

    def conv(name, kk, cin, cout):
        lim = (6.0 / (kk * kk * cin + kk * kk * cout)) ** 0.5
        w[name] = (rng.uniform(-lim, lim, (kk, kk, cin, cout)).astype(np.float32),
                   (0.05 * rng.standard_normal(cout) * affine_noise).astype(np.float32))

    def norm(name, c):
        w[name] = ((1 + 0.1 * rng.standard_normal(c) * affine_noise).astype(np.float32),
                   (0.1 * rng.standard_normal(c) * affine_noise).astype(np.float32))

    conv("conv3_u", 3, 128, 64); norm("norm3_u", 64)
    conv("conv2_u", 3, 64, 32); norm("norm2_u", 32)
    conv("conv1_u", 7, 32, 3)
    w["sun"] = random_sunrad_weights(seed + 99, H, W, affine_noise)
    return w


# ---------------------------------------------------------------------------------------------------------------------
# train / test step around the generator (train.py)
# ---------------------------------------------------------------------------------------------------------------------
def apply_rf(x, rf):
    """tf_utils.apply_rf / interp_1d / sample_1d (tf_utils.py:191-255): x [b, ...] in [0, 1], rf [b, k]."""
    b, k = rf.shape
    pos = (k - 1) * x.reshape(b, -1)
    y0 = torch.floor(pos)
    y1 = y0 + 1
    v0 = torch.gather(rf, 1, y0.long().clamp(0, k - 1))
    v1 = torch.gather(rf, 1, y1.long().clamp(0, k - 1))
    return ((y1 - pos) * v0 + (pos - y0) * v1).reshape(x.shape)


def ldr_synth(hdr, t, crf, sigma_s=None, sigma_c=None, noise_s=None, noise_c=None, quantize=True):
    """train._preprocessing (train.py:54-94) with the random draws as inputs; JPEG round trip omitted."""
    b = hdr.shape[0]
    x = hdr * t.reshape(b, 1, 1, 1)
    tmp = x
    if noise_s is not None:
        tmp = tmp + noise_s * (sigma_s.reshape(b, 1, 1, -1) * x)
    if noise_c is not None:
        tmp = tmp + sigma_c.reshape(b, 1, 1, -1) * noise_c
    hdr_t = torch.relu(tmp)
    ldr = apply_rf(hdr_t.clamp(0, 1), crf)
    if quantize:
        ldr = torch.round(ldr * 255.0) / 255.0
    return hdr_t, ldr


def kl_divergence(y_true, y_pred):
    """tf.keras.losses.KLDivergence: clip to [1e-7, 1], sum over the last axis, mean over the batch."""
    t, p = y_true.clamp(1e-7, 1.0), y_pred.clamp(1e-7, 1.0)
    return (t * torch.log(t / p)).sum(-1).mean()


def gaussian_filter2d(x, sigma):
    """tfa.image.gaussian_filter2d(filter_shape=(3, 3), sigma, padding='REFLECT') on NHWC."""
    g = torch.softmax(-torch.tensor([1.0, 0.0, 1.0], dtype=x.dtype) / (2.0 * sigma * sigma), 0)
    k2 = torch.outer(g, g)
    C = x.shape[-1]
    xp = torch.nn.functional.pad(x.permute(0, 3, 1, 2), (1, 1, 1, 1), mode="reflect")
    y = torch.nn.functional.conv2d(xp, k2[None, None].repeat(C, 1, 1, 1), groups=C)
    return y.permute(0, 2, 3, 1)


def dog(img):
    """tf_utils.DoG (tf_utils.py:61-73) -> four difference images."""
    B, h, w, C = img.shape
    up = O.resize_bilinear(img, 2 * h, 2 * w)
    base = gaussian_filter2d(up, 1.2489996)
    s1 = [1.2262735, 1.5450078, 1.9465878, 2.452547]
    s2 = [1.5450078, 1.9465878, 2.452547, 3.0900156]
    return [gaussian_filter2d(base, b) - gaussian_filter2d(base, a) for a, b in zip(s1, s2)]


def dog_l1(a, b):
    return sum((x - y).abs().mean() for x, y in zip(dog(a), dog(b)))


def vgg16_pools(bgr, data_dict, mean=(103.939, 116.779, 123.68), acc_dtype=torch.float32):
    """Vgg16.call (vgg16.py:130-165) -> (pool1, pool2, pool3)."""
    dt = acc_dtype
    x = 255.0 * bgr.to(dt) - torch.tensor(mean, dtype=dt)

    def conv(x, name):
        w, b = data_dict[name]
        return torch.relu(conv2d_same(x, w, b, acc_dtype=dt))

    p1 = maxpool2x2_same(conv(conv(x, "conv1_1"), "conv1_2"))
    p2 = maxpool2x2_same(conv(conv(p1, "conv2_1"), "conv2_2"))
    p3 = maxpool2x2_same(conv(conv(conv(p2, "conv3_1"), "conv3_2"), "conv3_3"))
    return p1, p2, p3


def discriminator(ldr, hdr, w, acc_dtype=torch.float32, training=False, bn_stats=None):
    """discriminator.model.call (discriminator.py:41-50); Conv2D(1, 4) is VALID.  training: see down_blocks."""
    dt = acc_dtype
    h = down_blocks(torch.cat([ldr.to(dt), hdr.to(dt)], dim=-1), w, dt, training, bn_stats)
    kern, bias = O._as_t(w["out"][0]).to(dt), O._as_t(w["out"][1]).to(dt)
    y = emu.conv2d(h.permute(0, 3, 1, 2), kern.permute(3, 2, 0, 1)) + bias.reshape(1, -1, 1, 1)
    return y.permute(0, 2, 3, 1)


def random_discriminator_weights(seed, bn_noise=True):
    w = random_sunrad_weights(seed, 32, 128, bn_noise)
    rng = np.random.default_rng(seed + 1)
    out = {k: w[k] for k in ("d1", "d2", "d3", "d4")}
    out["out"] = ((0.02 * rng.standard_normal((4, 4, 512, 1))).astype(np.float32), (0.1 * rng.standard_normal(1) * bn_noise).astype(np.float32))
    return out


def generator_test_step(ldr, hdr_t, sunpose_gt, wg, ws, wd, vgg_dict, k=3, threshold=0.12, acc_dtype=torch.float32):
    """train.generator_in_step(training=False) + discriminator_in_step(training=False) (train.py:239-380) -> dict of losses."""
    dt = acc_dtype
    inp = O._as_t(ldr).to(dt)
    hdr_t = O._as_t(hdr_t).to(dt)
    gt = O._as_t(sunpose_gt).to(dt)
    B, H, W, _ = inp.shape

    def norm_act(y, name):
        g, b = (O._as_t(v).to(dt) for v in wg[name])
        return leaky_relu(instance_norm(y, g, b), 0.1)

    x = norm_act(conv2d_same(inp, *wg["conv1_d"], stride=1, acc_dtype=dt), "norm1_d")
    x = norm_act(conv2d_same(x, *wg["conv2_d"], stride=2, acc_dtype=dt), "norm2_d")
    x = norm_act(conv2d_same(x, *wg["conv3_d"], stride=2, acc_dtype=dt), "norm3_d")
    blocks = [{kk: O._as_t(v) for kk, v in blk.items()} for blk in wg["res"]]
    res_out = res_layer(x, blocks, k, acc_dtype=dt)
    sky_gamma = decode_branch(res_out, inp, wg, ("conv3_f", "norm3_f", "conv2_f", "norm2_f", "conv1_f"), dt)
    sky_lin = hdr_log_decompression(sky_gamma)
    x_sun = inp.clone().requires_grad_(True)
    sm, acts = sunpose_estimation(x_sun, ws, distortion_aware=True, acc_dtype=dt)
    alpha = torch.clamp((sky_lin.amax(dim=3) - 1.0 + threshold).clamp(min=0.0) / threshold, max=1.0).unsqueeze(-1)
    y_c = torch.gather(sm, 1, gt.argmax(dim=1, keepdim=True))[:, 0]                     # train.py:278-280
    cams = [grad_cam_layer(y_c, a).detach() for a in acts]
    sm = sm.detach()
    pred = sm.reshape(B, H, W, 1)
    plz = torch.cat([inp, cams[0], O.resize_bilinear(cams[1], H, W), O.resize_bilinear(cams[2], H, W)], dim=-1)
    sun_rad = sunrad_net(pred / pred.max(), plz, wg["sun"], acc_dtype=dt).expand(B, H, W, 3)
    sun_gamma = decode_branch(res_out, hdr_log_compression(sun_rad), wg, ("conv3_u", "norm3_u", "conv2_u", "norm2_u", "conv1_u"), dt)
    sky_s, sun_s = (1.0 - alpha) * sky_gamma, alpha * sun_gamma
    y_gamma = sky_s + sun_s
    y_lin = hdr_log_decompression(y_gamma)
    out = dict(y_final_gamma=y_gamma, y_final_lin=y_lin, sky_pred_lin=hdr_log_decompression(sky_s), sun_pred_lin=hdr_log_decompression(sun_s))
    out["kl"] = kl_divergence(gt, sm)
    pa, pb = vgg16_pools(y_gamma, vgg_dict, acc_dtype=dt), vgg16_pools(hdr_log_compression(hdr_t), vgg_dict, acc_dtype=dt)
    out["perceptual"] = sum((a - b).abs().mean() for a, b in zip(pa, pb))
    out["dog"] = dog_l1(y_lin, hdr_t)
    out["l1"] = (y_lin - hdr_t).abs().mean()
    d_fake = discriminator(inp, y_lin, wd, dt)
    out["gen"] = ((d_fake - 1.0) ** 2).mean()
    out["total"] = out["kl"] + 1000.0 * out["dog"] + out["gen"] + 10.0 * out["l1"] + 0.01 * out["perceptual"]
    d_real = discriminator(inp, hdr_t, wd, dt)
    out["disc"] = 0.5 * ((d_fake ** 2).mean() + ((d_real - 1.0) ** 2).mean())
    return out


def sun_train_step_grads(ldr, sunpose_gt, ws, acc_dtype=torch.float64, with_gradcam=False):
    """train_sun.sun_train_step (train_sun.py:220-264) up to the gradients: loss = KLDivergence(gt, sm) + sum_l mean|DoG_l(pred) - DoG_l(gt)|,
    autograd w.r.t. every variable of the sun-position network (distortion-aware wiring).  Returns (loss, grads in the layout of ws)."""
    dt = acc_dtype
    leaves = {}

    def leaf(a):
        t = O._as_t(a).to(dt).clone().requires_grad_(True)
        return t

    w = {}
    for name in ("sunlayer1", "sunlayer2", "sunlayer3"):
        w[name] = {k: leaf(v) for k, v in ws[name].items()}
    for name in ("fc1", "fc2"):
        w[name] = tuple(leaf(v) for v in ws[name])
    x = O._as_t(ldr).to(dt)
    gt = O._as_t(sunpose_gt).to(dt)
    B, H, W, _ = x.shape
    if with_gradcam:
        x = x.clone().requires_grad_(True)
    sm, acts = sunpose_estimation(x, w, distortion_aware=True, acc_dtype=dt)
    if with_gradcam:        # train_sun.py:231-240 (outside the tape: outputs only)
        y_c = torch.gather(sm, 1, gt.argmax(dim=1, keepdim=True))[:, 0]
        cams = [grad_cam_layer(y_c, a).detach() for a in acts]
    loss = kl_divergence(gt, sm) + dog_l1(sm.reshape(B, H, W, 1), gt.reshape(B, H, W, 1))
    flat = [w[n][k] for n in ("sunlayer1", "sunlayer2", "sunlayer3") for k in sorted(w[n])] + [t for n in ("fc1", "fc2") for t in w[n]]
    grads = torch.autograd.grad(loss, flat)
    it = iter(grads)
    out = {n: {k: next(it) for k in sorted(w[n])} for n in ("sunlayer1", "sunlayer2", "sunlayer3")}
    for n in ("fc1", "fc2"):
        out[n] = (next(it), next(it))
    return loss.detach(), out


# ---------------------------------------------------------------------------------------------------------------------
# full train step (train.py:382-415): both tapes
# ---------------------------------------------------------------------------------------------------------------------
def _leafify(obj, dt):
    """Nested dict / list / tuple of arrays -> the same nesting of fp leaves that require grad."""
    if isinstance(obj, dict):
        return {k: _leafify(v, dt) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_leafify(v, dt) for v in obj)
    return O._as_t(obj).to(dt).clone().requires_grad_(True)


def _flatten(obj, out, prefix=""):
    if isinstance(obj, dict):
        for k in obj:
            _flatten(obj[k], out, f"{prefix}{k}.")
    elif isinstance(obj, (list, tuple)):
        for i, v in enumerate(obj):
            _flatten(v, out, f"{prefix}{i}.")
    else:
        out.append((prefix[:-1], obj))
    return out


def generator_train_forward(inp, hdr_t, gt, wg, ws, wd, vgg_dict, k=3, threshold=0.12, dt=torch.float64, distortion_aware_sunpose=True,
                            bn_stats=None):
    """train.generator_in_step(training=True) (train.py:239-349) on autograd leaves: alpha and the Grad-CAM maps are computed under
    gen_tape.stop_recording() (:256-271) and are constants; sunRadNet runs with batch statistics (sun_rad_estimation passes `training`
    through, generator.py:165); the discriminator call inside the generator step uses its moving statistics (train.py:300)."""
    B, H, W, _ = inp.shape

    def norm_act(y, name):
        return leaky_relu(instance_norm(y, wg[name][0], wg[name][1]), 0.1)

    x = norm_act(conv2d_same(inp, *wg["conv1_d"], stride=1, acc_dtype=dt), "norm1_d")
    x = norm_act(conv2d_same(x, *wg["conv2_d"], stride=2, acc_dtype=dt), "norm2_d")
    x = norm_act(conv2d_same(x, *wg["conv3_d"], stride=2, acc_dtype=dt), "norm3_d")
    res_out = res_layer(x, wg["res"], k, acc_dtype=dt)
    sky_gamma = decode_branch(res_out, inp, wg, ("conv3_f", "norm3_f", "conv2_f", "norm2_f", "conv1_f"), dt)
    sky_lin = hdr_log_decompression(sky_gamma)
    sm, acts = sunpose_estimation(inp, ws, distortion_aware=distortion_aware_sunpose, acc_dtype=dt)
    pred = sm.reshape(B, H, W, 1)
    with torch.no_grad():
        alpha = torch.clamp((sky_lin.amax(dim=3) - 1.0 + threshold).clamp(min=0.0) / threshold, max=1.0).unsqueeze(-1)
    y_c = torch.gather(sm, 1, gt.argmax(dim=1, keepdim=True))[:, 0]                     # train.py:263-265
    cams = [grad_cam_layer(y_c, a).detach() for a in acts]                                # :267-269, outside the tape
    plz = torch.cat([inp, cams[0], O.resize_bilinear(cams[1], H, W), O.resize_bilinear(cams[2], H, W)], dim=-1)
    normed = pred / pred.max()                                                            # generator.py:160 (differentiable, max included)
    sun_rad, gamma_in, beta_in = sunrad_net(normed, plz, wg["sun"], acc_dtype=dt, return_heads=True, training=True, bn_stats=bn_stats)
    sun_rad_gamma = hdr_log_compression(sun_rad.expand(B, H, W, 3))
    sun_gamma = decode_branch(res_out, sun_rad_gamma, wg, ("conv3_u", "norm3_u", "conv2_u", "norm2_u", "conv1_u"), dt)
    sky_s, sun_s = (1.0 - alpha) * sky_gamma, alpha * sun_gamma
    y_gamma = sky_s + sun_s
    y_lin = hdr_log_decompression(y_gamma)
    out = dict(y_final_gamma=y_gamma, y_final_lin=y_lin, sky_pred_lin=hdr_log_decompression(sky_s), sun_pred_lin=hdr_log_decompression(sun_s),
               alpha=alpha, sm=sm, cams=cams, sun_rad_lin=sun_rad, gamma=gamma_in, beta=beta_in)
    out["kl"] = kl_divergence(gt, sm)
    pa = vgg16_pools(y_gamma, vgg_dict, acc_dtype=dt)
    with torch.no_grad():
        pb = vgg16_pools(hdr_log_compression(hdr_t), vgg_dict, acc_dtype=dt)
    out["perceptual"] = sum((a - b).abs().mean() for a, b in zip(pa, pb))
    out["dog"] = dog_l1(y_lin, hdr_t)
    out["l1"] = (y_lin - hdr_t).abs().mean()
    d_fake = discriminator(inp, y_lin, wd, dt, training=False)
    out["gen"] = ((d_fake - 1.0) ** 2).mean()
    out["total"] = out["kl"] + 1000.0 * out["dog"] + out["gen"] + 10.0 * out["l1"] + 0.01 * out["perceptual"]
    return out


def train_step(ldr, hdr_t, sunpose_gt, wg, ws, wd, vgg_dict, k=3, threshold=0.12, acc_dtype=torch.float64, distortion_aware_sunpose=True):
    """train.train_step (train.py:382-415) up to the optimizer: total_gen_loss differentiated w.r.t. every variable of _gen and _sun
    (:402), total_disc_loss (discriminator_in_step, :351-380, BatchNormalization with batch statistics, one set per call) w.r.t. every
    variable of _dis (:405).  Returns dict(losses..., outputs..., grads_gen, grads_sun, grads_dis (nested like the weight dicts, None for
    the non-trainable moving statistics), bn=dict(sun / dis_real / dis_fake -> name -> (mean, biased var, count)))."""
    dt = acc_dtype
    Lg, Ls, Ld = _leafify(wg, dt), _leafify(ws, dt), _leafify(wd, dt)
    inp, hdr_t, gt = O._as_t(ldr).to(dt), O._as_t(hdr_t).to(dt), O._as_t(sunpose_gt).to(dt)
    bn = dict(sun={}, dis_real={}, dis_fake={})
    out = generator_train_forward(inp, hdr_t, gt, Lg, Ls, Ld, vgg_dict, k, threshold, dt, distortion_aware_sunpose, bn["sun"])
    y_lin = out["y_final_lin"].detach()                                                   # D's tape never reaches the generator variables
    d_real = discriminator(inp, hdr_t, Ld, dt, training=True, bn_stats=bn["dis_real"])    # train.py:360
    d_fake = discriminator(inp, y_lin, Ld, dt, training=True, bn_stats=bn["dis_fake"])    # :361
    out["disc"] = 0.5 * ((d_fake ** 2).mean() + ((d_real - 1.0) ** 2).mean())             # :364-368
    gen_leaves = _flatten(Lg, []) + [("sun." + n, t) for n, t in _flatten(Ls, [])]
    g = torch.autograd.grad(out["total"], [t for _, t in gen_leaves], retain_graph=True, allow_unused=True)
    dis_leaves = _flatten(Ld, [])
    gd = torch.autograd.grad(out["disc"], [t for _, t in dis_leaves], allow_unused=True)
    res = {kk: (v.detach() if isinstance(v, torch.Tensor) else [c.detach() for c in v]) for kk, v in out.items()}
    res["grads_gen"] = {n: gg for (n, _), gg in zip(gen_leaves, g)}
    res["grads_dis"] = {n: gg for (n, _), gg in zip(dis_leaves, gd)}
    res["bn"] = bn
    return res


def rmsprop_step(w, ms, g, lr=1e-4, rho=0.9, eps=1e-7):
    """Keras RMSprop (train.py:201-202): ms = rho ms + (1 - rho) g^2; w -= lr g / (sqrt(ms) + eps)."""
    ms = rho * ms + (1.0 - rho) * g * g
    return w - lr * g / (np.sqrt(ms) + eps), ms
