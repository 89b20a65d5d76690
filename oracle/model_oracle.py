"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/da_oracle.py for the rules).

CPU (torch fp32 / fp64) restatement of the residual trunk of the reference generator with the distortion-aware convs
wired in as the commented lines generator.py:14,18 do:

  instance_norm()   tfa.layers.InstanceNormalization defaults (generator.py:15,19): per (n, c) over H x W, biased variance
                    (tf.nn.moments), eps 1e-3, y = x * (gamma * rsqrt(var + eps)) + (beta - mean * gamma * rsqrt(var + eps))
                    — TensorFlow-Addons is not vendored in the reference; restated from its published GroupNormalization code
  res_block()       generator.py:26-35
  res_layer()       generator.py:46-49
"""
from __future__ import annotations

import torch

from . import da_oracle as O


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
    return x.to(dt) + n2


def res_layer(x, blocks, k=3, dilation=1, acc_dtype=torch.float32):
    x = O._as_t(x).to(acc_dtype)
    for w in blocks:
        x = res_block(x, w, k, dilation, acc_dtype)
    return x


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
