"""GPU parity tests of the distortion-aware conv path, through the C ABI (ctypes layer shim), against the oracle and the
golden vectors.  Bar: offsets / indices / bilinear weights bit-exact; layer outputs within the stated tolerances:

  TF32 tensor-core path   relative L2 <= 1.5e-3 per layer vs the fp64-accumulated oracle (operands rounded to 10-bit mantissa)
  3xTF32 path / SIMT      relative L2 <= 2e-5 / 1e-5
"""
import numpy as np
import pytest
import torch

from conftest import golden_cases
from oracle import da_oracle as O

pytestmark = pytest.mark.gpu

TOL_TF32, TOL_3X, TOL_SIMT = 1.5e-3, 2e-5, 1e-5


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.int32)


@pytest.fixture(scope="module")
def ops(pkg):
    assert torch.cuda.is_available()
    return pkg.distortion_aware_ops


def test_device_offsets_kernel(ops):
    for (h, w, k, dil, sky) in [(8, 32, 3, 1, True), (32, 128, 3, 1, True), (32, 128, 7, 1, True), (64, 256, 3, 1, False),
                                (128, 512, 3, 1, True), (128, 512, 7, 1, True), (32, 128, 3, 2, True)]:
        dev = ops.offsets_table(h, w, k, dil, sky, mode="device").cpu().numpy()
        # bit-exact with the correctly-rounded flavour of the oracle ...
        assert np.array_equal(bits(dev), bits(O.offsets(h, w, k, dil, sky, via_f64=True))), (h, w, k)
        # ... and within a few ulp of glibc's float libm (~5% of entries differ in the last bits)
        ref = O.offsets(h, w, k, dil, sky)
        assert np.abs(dev - ref).max() <= 8 * np.spacing(np.float32(np.abs(ref).max()))
        # the default (libm) mode is bit-exact with the oracle
        assert np.array_equal(bits(ops.offsets_table(h, w, k, dil, sky).cpu().numpy()), bits(ref))
    with pytest.raises(Exception, match="undefined coordinates"):
        ops.offsets_table(2, 8, 3, mode="device")


def test_sampling_indices_and_weights_bit_exact(ops, golden):
    for (h, w, k) in [(8, 32, 3), (32, 128, 3), (16, 64, 7), (64, 256, 3), (32, 128, 7), (12, 48, 5)]:
        tab = ops.offsets_table(h, w, k)
        got = {n: t.cpu().numpy() for n, t in ops.sample_debug(h, w, k, tab).items()}
        want = O.sample(h, w, k, O.offsets(h, w, k))
        for n in ("y0", "y1", "x0", "x1"):
            assert np.array_equal(got[n], want[n]), (h, w, k, n)
        for n in ("w0", "w1", "w2", "w3"):
            assert np.array_equal(bits(got[n]), bits(want[n])), (h, w, k, n)
    for c in golden_cases(golden):          # and directly against what the reference source fed to tf.gather_nd
        h, w = (c["h"], c["w"]) if c["kind"] == "conv" else c["out_hw"]
        tab = ops.offsets_table(h, w, c["k"], c["dilation"], c["skydome"])
        got = {n: t.cpu().numpy() for n, t in ops.sample_debug(h, w, c["k"], tab).items()}
        idx, wg = golden[f"idx__{c['name']}"], golden[f"wgt__{c['name']}"]
        for ci, (yy, xx) in enumerate((("y0", "x0"), ("y0", "x1"), ("y1", "x0"), ("y1", "x1"))):
            assert np.array_equal(got[yy], idx[ci, ..., 0]) and np.array_equal(got[xx], idx[ci, ..., 1]), c["name"]
            assert np.array_equal(bits(got[f"w{ci}"]), bits(wg[ci])), c["name"]


def _layer(pkg, kind, c, kern, b, math_mode):
    if kind == "conv":
        return pkg.conv2d(c["F"], kernel_size=c["k"], strides=1, dilation_rate=c["dilation"], skydome=c["skydome"],
                          kernel_initializer=kern, bias_initializer=b, math_mode=math_mode)
    return pkg.deconv2d(c["F"], kernel_size=c["k"], output_imshape=list(c["out_hw"]), dilation_rate=c["dilation"],
                        skydome=c["skydome"], kernel_initializer=kern, bias_initializer=b, math_mode=math_mode)


def test_golden_layer_outputs(pkg, ops, golden):
    for c in golden_cases(golden):
        name = c["name"]
        x, kern, b, want = (golden[f"{k}__{name}"] for k in ("x", "k", "b", "y"))
        xd = torch.from_numpy(x).cuda()
        for mode, tol in (("tf32", TOL_TF32), ("3xtf32", TOL_3X)):
            got = _layer(pkg, c["kind"], c, kern, b, mode)(xd).cpu().numpy()
            assert got.shape == want.shape
            assert rel_l2(got, want) <= tol, (name, mode, rel_l2(got, want))
        if c["kind"] == "deconv":
            r = ops.resize_bilinear(xd, *c["out_hw"]).cpu().numpy()
            assert np.array_equal(bits(r), bits(golden[f"resized__{name}"])), name


CASES = [
    # B, h, w, C, F, k      what it exercises
    (2, 8, 32, 128, 128, 3),    # res-trunk site at 32x128 (generator.py:13-19), M = 512
    (3, 8, 32, 64, 96, 3),      # M = 768, F not a power of two
    (1, 5, 24, 32, 16, 3),      # M = 120 < one tile (ragged tail), odd h
    (2, 16, 64, 32, 32, 7),     # sunposeLayer 7x7 site (sunpose_net.py:36-40)
    (2, 16, 64, 3, 32, 7),      # 3-channel image input: generic-C producer path, K = 147 padded to 160
    (2, 16, 64, 32, 3, 7),      # 3 filters: N padded to 16, scalar stores
    (1, 8, 32, 40, 24, 3),      # C % 32 != 0 and C % 4 == 0
    (1, 8, 32, 32, 256, 3),     # widest N
    (1, 32, 128, 64, 64, 3),    # full-resolution map
    (2, 64, 256, 32, 32, 3),    # wide map: tiles away from the seam, tiles on it
    (1, 12, 48, 32, 16, 5),     # k=5, sizes that are not multiples of the tile
    (5, 4, 16, 64, 32, 3),      # map smaller than a tile (64 pixels per panorama)
]


@pytest.mark.parametrize("B,h,w,C,F,k", CASES)
def test_conv_forward_vs_oracle(pkg, ops, B, h, w, C, F, k):
    rng = np.random.default_rng(B * 1000 + h * 10 + C + F + k)
    x = rng.standard_normal((B, h, w, C)).astype(np.float32)
    kern = (rng.standard_normal((k * k * C, F)) / np.sqrt(k * k * C)).astype(np.float32)
    b = rng.standard_normal(F).astype(np.float32)
    want = O.conv2d_forward(x, kern, b, k, acc_dtype=torch.float64).numpy()
    xd = torch.from_numpy(x).cuda()
    tab = ops.offsets_table(h, w, k)
    simt = ops.conv2d_simt(xd, tab, torch.from_numpy(kern).cuda(), torch.from_numpy(b).cuda(), k).cpu().numpy()
    assert rel_l2(simt, want) <= TOL_SIMT, rel_l2(simt, want)
    for mode, tol in (("tf32", TOL_TF32), ("3xtf32", TOL_3X)):
        layer = pkg.conv2d(F, kernel_size=k, kernel_initializer=kern, bias_initializer=b, math_mode=mode)
        got = layer(xd).cpu().numpy()
        assert np.isfinite(got).all()
        assert rel_l2(got, want) <= tol, (mode, rel_l2(got, want))
        # the default is the row-strip kernel where it applies (C % 32 == 0); the band-staged producer kernel of round 1 and the
        # direct-gather kernel (no smem band) must agree with it
        band = layer(xd, kernel_path="band").cpu().numpy()
        assert rel_l2(band, want) <= tol, (mode, "band", rel_l2(band, want))
        direct = layer(xd, force_direct=True).cpu().numpy()
        assert rel_l2(direct, want) <= tol, (mode, "direct", rel_l2(direct, want))
        # fused instance-norm moments: sum and sum of squares of y per (sample, filter)
        stats = torch.zeros(B, F, 2, dtype=torch.float64, device="cuda")
        y2 = layer(xd, stats=stats)
        ref_s = torch.stack([y2.double().sum((1, 2)), (y2.double() ** 2).sum((1, 2))], dim=-1)
        assert torch.allclose(stats, ref_s, rtol=1e-5, atol=1e-4), (stats - ref_s).abs().max()
        assert tuple(layer.offset.shape) == (1, h, w, k * k, 2)
        assert tuple(layer.kernel.shape) == (k * k * C, F) and tuple(layer.bias.shape) == (F,)


def test_epilogue_leaky_relu_and_residual(pkg):
    rng = np.random.default_rng(7)
    B, h, w, C, F, k = 2, 8, 32, 32, 32, 3
    x = rng.standard_normal((B, h, w, C)).astype(np.float32)
    kern = (rng.standard_normal((k * k * C, F)) / 17).astype(np.float32)
    b = rng.standard_normal(F).astype(np.float32)
    res = rng.standard_normal((B, h, w, F)).astype(np.float32)
    base = O.conv2d_forward(x, kern, b, k, acc_dtype=torch.float64).numpy()
    want = np.where(base > 0, base, 0.1 * base) + res
    layer = pkg.conv2d(F, kernel_size=k, kernel_initializer=kern, bias_initializer=b, math_mode="3xtf32")
    got = layer(torch.from_numpy(x).cuda(), leaky_slope=0.1, residual=torch.from_numpy(res).cuda()).cpu().numpy()
    assert rel_l2(got, want) <= TOL_3X


def test_deconv_forward_vs_oracle(pkg):
    rng = np.random.default_rng(11)
    B, h, w, C, F, k = 2, 8, 32, 128, 64, 3          # sky_decode's first resize-deconv site (generator.py:112)
    x = rng.standard_normal((B, h, w, C)).astype(np.float32)
    kern = (rng.standard_normal((k * k * C, F)) / np.sqrt(k * k * C)).astype(np.float32)
    b = rng.standard_normal(F).astype(np.float32)
    want = O.deconv2d_forward(x, kern, b, (16, 64), k, acc_dtype=torch.float64).numpy()
    for mode, tol in (("tf32", TOL_TF32), ("3xtf32", TOL_3X)):
        layer = pkg.deconv2d(F, kernel_size=k, output_imshape=[16, 64], kernel_initializer=kern, bias_initializer=b,
                             math_mode=mode)
        got = layer(torch.from_numpy(x).cuda()).cpu().numpy()
        assert got.shape == (B, 16, 64, F)
        assert rel_l2(got, want) <= tol, (mode, rel_l2(got, want))


def test_reference_restrictions(pkg):
    x = torch.zeros(1, 8, 32, 4, device="cuda")
    with pytest.raises(ValueError, match="strides=1"):
        pkg.conv2d(4, strides=2)(x)
    with pytest.raises(AssertionError, match="kernel_size must be odd"):
        pkg.conv2d(4, kernel_size=4)(x)
    with pytest.raises(RuntimeError, match="no CPU path"):
        pkg.conv2d(4)(x.cpu())
    layer = pkg.conv2d(4)
    layer(x)
    with pytest.raises(ValueError, match="static shapes"):
        layer(torch.zeros(1, 16, 32, 4, device="cuda"))


def test_full_size_properties(pkg, ops):
    """BASELINE config 4 size (B=64, 64x256, 128->128, k=3): properties that need no CPU oracle run."""
    torch.manual_seed(0)
    B, h, w, C, F, k = 64, 64, 256, 128, 128, 3
    x = torch.randn(B, h, w, C, device="cuda")
    layer = pkg.conv2d(F, kernel_size=k, math_mode="tf32")
    layer.build(tuple(x.shape))
    layer.bias.normal_()
    y = layer(x)
    # (1) batch-permutation equivariance, bit for bit (tiles never straddle samples: h*w % 128 == 0)
    perm = torch.randperm(B, device="cuda")
    assert torch.equal(layer(x[perm].contiguous()), y[perm])
    # (2) zero input -> bias exactly
    z = layer(torch.zeros_like(x))
    assert torch.equal(z, layer.bias.expand_as(z))
    # (3) tensor-core vs CUDA-core restatement on a slice of samples
    simt = ops.conv2d_simt(x[:2].contiguous(), layer.offset_table, layer.kernel, layer.bias, k)
    r = (torch.linalg.norm((y[:2] - simt).double()) / torch.linalg.norm(simt.double())).item()
    assert r <= TOL_TF32, r
    # (4) linearity in x within the TF32 tolerance
    x2 = torch.randn_like(x[:4])
    lhs = layer((x[:4] * 0.5 + x2).contiguous()) - layer.bias
    rhs = 0.5 * (y[:4] - layer.bias) + (layer(x2) - layer.bias)
    r = (torch.linalg.norm((lhs - rhs).double()) / torch.linalg.norm(rhs.double())).item()
    assert r <= 2 * TOL_TF32, r


def test_tile_pairs_equal_single_tiles(pkg):
    """Large launches process two tiles of a row class per weight tile (M = 256 through two accumulators).  Per tile the MMA sequence is
    the same as without pairing, so the results are bit-identical — also when a row class has an odd number of tiles (the last pair's
    second accumulator is never stored)."""
    torch.manual_seed(1)
    for (B, h, w, C, F) in ((3, 150, 128, 32, 32), (32, 32, 128, 64, 24)):
        x = torch.randn(B, h, w, C, device="cuda")
        layer = pkg.conv2d(F, kernel_size=3, math_mode="tf32")
        layer.build(tuple(x.shape))
        layer.bias.normal_()
        stats_a = torch.zeros(B, F, 2, dtype=torch.float64, device="cuda")
        stats_b = torch.zeros_like(stats_a)
        paired = layer(x, stats=stats_a)
        single = layer(x, stats=stats_b, extra_flags=pkg._lib.EPI_NO_PAIR)
        assert torch.equal(paired, single)
        assert torch.allclose(stats_a, stats_b, rtol=1e-12, atol=1e-9)
