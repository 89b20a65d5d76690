"""GPU parity of the plain encoder/decoder convs (ops.py) and of the sky branch of generator inference
(inference.py:84-86 = encode -> sky_decode -> hdr_logDecompression) against the oracle.

Tolerances: per plain conv (TF32) relative L2 <= 1.5e-3; the whole sky branch in the LOG domain (sky_pred_gamma, what
north_star's bound is stated on) <= 5e-3 for TF32 and <= 3e-4 for 3xTF32; the linear radiance after exp() <= 2e-2 / 1e-3."""
import numpy as np
import pytest
import torch

from oracle import model_oracle as M

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize("B,h,w,C,F,k,s", [
    (2, 32, 128, 3, 32, 7, 1),      # conv1_d (generator.py:60)
    (2, 32, 128, 32, 64, 3, 2),     # conv2_d: stride 2, SAME pad 0 / 1
    (2, 16, 64, 64, 128, 3, 2),     # conv3_d
    (2, 16, 64, 128, 64, 3, 1),     # conv3_f after the resize
    (1, 32, 128, 32, 3, 7, 1),      # conv1_f
    (1, 9, 21, 32, 16, 3, 2),       # odd sizes: out = ceil(n/2), pad 1 / 1
    (1, 8, 32, 6, 64, 4, 2),        # sunRadNet's first layer shape (4x4 stride 2, even kernel)
])
def test_plain_conv_vs_oracle(pkg, B, h, w, C, F, k, s):
    rng = np.random.default_rng(h + C + F + k + s)
    x = rng.standard_normal((B, h, w, C)).astype(np.float32)
    w4 = (rng.standard_normal((k, k, C, F)) / np.sqrt(k * k * C)).astype(np.float32)
    b = rng.standard_normal(F).astype(np.float32)
    want = M.conv2d_same(x, w4, b, stride=s, acc_dtype=torch.float64).numpy()
    for mode, tol in (("tf32", 1.5e-3), ("3xtf32", 2e-5)):
        layer = pkg.ops.conv2d(output_channels=F, strides=s, k_h=k, k_w=k, kernel_initializer=w4, bias_initializer=b,
                               math_mode=mode)
        got = layer(torch.from_numpy(x).cuda()).cpu().numpy()
        assert got.shape == want.shape
        assert rel_l2(got, want) <= tol, (mode, rel_l2(got, want))


def test_conv_epilogue_tail(pkg):
    """leaky_relu(conv) + input -> relu -> log-decompress fused in the last conv (generator.py:120-124, tf_utils.py:273-280)."""
    rng = np.random.default_rng(2)
    B, h, w, C, F, k = 1, 16, 64, 32, 3, 7
    x = rng.standard_normal((B, h, w, C)).astype(np.float32)
    w4 = (rng.standard_normal((k, k, C, F)) / np.sqrt(k * k * C)).astype(np.float32)
    b = rng.standard_normal(F).astype(np.float32)
    inp = rng.uniform(0, 1, (B, h, w, F)).astype(np.float32)
    c = M.conv2d_same(x, w4, b, acc_dtype=torch.float64)
    want = M.hdr_log_decompression(torch.relu(torch.from_numpy(inp).double() + M.leaky_relu(c, 0.1))).numpy()
    layer = pkg.ops.conv2d(output_channels=F, strides=1, k_h=k, k_w=k, kernel_initializer=w4, bias_initializer=b, math_mode="3xtf32")
    got = layer(torch.from_numpy(x).cuda(), leaky_slope=0.1, residual=torch.from_numpy(inp).cuda(), relu=True,
                log_decompress=True).cpu().numpy()
    assert rel_l2(got, want) <= 1e-4, rel_l2(got, want)


@pytest.mark.parametrize("mode,tol_log,tol_lin", [("tf32", 5e-3, 2e-2), ("3xtf32", 3e-4, 1e-3)])
def test_sky_branch_inference_vs_oracle(pkg, mode, tol_log, tol_lin):
    rng = np.random.default_rng(0)
    B, H, W = 2, 32, 128
    ldr = (np.round(255 * rng.uniform(0, 1, (B, H, W, 3))) / 255).astype(np.float32)      # SURVEY 8d synthetic LDR
    w = M.random_generator_weights(seed=4)
    gen = pkg.model(batch_size=B, im_height=H, im_width=W, da_kernel_size=3, math_mode=mode)
    gen.build(B)
    gen.set_weights(w)
    x = torch.from_numpy(ldr).cuda()
    got_log = gen.sky_decode(gen.encode(x), x).cpu().numpy()
    got_lin = gen.sky_inference(x).cpu().numpy()
    want_log = M.sky_branch(ldr, w, acc_dtype=torch.float64, decompress=False).numpy()
    want_lin = M.sky_branch(ldr, w, acc_dtype=torch.float64, decompress=True).numpy()
    assert got_log.shape == (B, H, W, 3)
    assert rel_l2(got_log, want_log) <= tol_log, ("log", rel_l2(got_log, want_log))
    assert rel_l2(got_lin, want_lin) <= tol_lin, ("lin", rel_l2(got_lin, want_lin))


@pytest.mark.parametrize("C,F,k,shape", [(32, 3, 7, (2, 32, 128)), (32, 3, 3, (1, 13, 45)), (8, 1, 5, (2, 9, 33)), (64, 4, 3, (1, 16, 64))])
def test_small_filter_count_conv(pkg, C, F, k, shape):
    """Plain stride-1 layers with F <= 4 filters (conv1_f / conv1_u) take the fp32 CUDA-core kernel that reads the TF32-rounded packed
    weights: ragged tiles, C % 32 != 0 packing order, every epilogue flag.  2e-3 relative L2 after the exponential of the decompression (weights rounded to TF32, fp32 FMA)."""
    rng = np.random.default_rng(C + F + k)
    B, h, w = shape
    x = rng.standard_normal((B, h, w, C)).astype(np.float32)
    kern = (rng.standard_normal((k, k, C, F)) / np.sqrt(k * k * C)).astype(np.float32)
    bias = rng.standard_normal(F).astype(np.float32)
    res = rng.uniform(0, 1, (B, h, w, F)).astype(np.float32)
    layer = pkg.ops.conv2d(output_channels=F, k_h=k, k_w=k, strides=1, kernel_initializer=kern, bias_initializer=bias)
    got = layer(torch.from_numpy(x).cuda(), leaky_slope=0.1, residual=torch.from_numpy(res).cuda(), relu=True, log_decompress=True).cpu().numpy()
    y = M.conv2d_same(x, kern, bias, acc_dtype=torch.float64)
    want = M.hdr_log_decompression(torch.relu(M.leaky_relu(y, 0.1) + torch.from_numpy(res).double())).numpy()
    r = np.linalg.norm(got - want) / np.linalg.norm(want)
    assert r < 2e-3, r
