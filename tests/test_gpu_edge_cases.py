"""Edge cases of the sun-branch / train-tail kernels through the C ABI: boundary values, minimal sizes, ties, ragged shapes, and the
consistency of the fused and un-fused tails."""
import numpy as np
import pytest
import torch

from oracle import model_oracle as M

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _st():
    return torch.cuda.current_stream().cuda_stream


def test_ldr_synth_boundaries(pkg):
    # x = 0, x = 1 exactly (pos = K-1: the upper neighbour index K is clamped), x > 1 (clipped), K = 2, no noise inputs
    for K in (2, 5, 1024):
        crf = np.stack([np.linspace(0, 1, K) ** 0.5, np.linspace(0, 1, K)]).astype(np.float32)
        hdr = np.array([0.0, 1.0, 0.999999, 2.5, 1e-8, 0.5], np.float32).reshape(1, 1, 6, 1).repeat(2, 0)
        t = np.array([1.0, 2.0], np.float32)
        T = torch.from_numpy
        got_h, got_l = pkg.tf_utils.ldr_synth(T(hdr).cuda(), T(t).cuda(), T(crf).cuda(), quantize=False)
        want_h, want_l = M.ldr_synth(T(hdr).double(), T(t).double(), T(crf).double(), quantize=False)
        assert np.abs(got_h.cpu().numpy() - want_h.numpy()).max() < 1e-6
        assert np.abs(got_l.cpu().numpy() - want_l.numpy()).max() < 2e-6
        assert float(got_l.max()) <= 1.0 and float(got_l.min()) >= 0.0


def test_dog_minimal_image_and_identical_inputs(pkg):
    rng = np.random.default_rng(0)
    a = rng.uniform(0, 1, (1, 2, 2, 1)).astype(np.float32)
    b = rng.uniform(0, 1, (1, 2, 2, 1)).astype(np.float32)
    acc = torch.zeros(4, dtype=torch.float64, device="cuda")
    got = float(pkg.tf_utils.DoG_l1(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), acc))
    want = float(M.dog_l1(torch.from_numpy(a).double(), torch.from_numpy(b).double()))
    assert abs(got / want - 1) < 1e-4
    acc.zero_()
    assert float(pkg.tf_utils.DoG_l1(torch.from_numpy(a).cuda(), torch.from_numpy(a).cuda(), acc)) == 0.0


@pytest.mark.parametrize("C", [32, 64, 128])
def test_gradcam_kernel(pkg, C):
    rng = np.random.default_rng(C)
    B, h, w = 3, 5, 12
    g = rng.standard_normal((B, h, w, C)).astype(np.float32)
    A = np.maximum(rng.standard_normal((B, h, w, C)), 0).astype(np.float32)
    wsum = torch.empty((B, C), device="cuda")
    cam = torch.empty((B, h, w), device="cuda")
    gd, Ad = torch.from_numpy(g).cuda(), torch.from_numpy(A).cuda()
    pkg._lib.check(pkg._lib.LIB.sky_gradcam(gd.data_ptr(), Ad.data_ptr(), wsum.data_ptr(), cam.data_ptr(), B, h, w, C, _st()))
    want = np.maximum(np.einsum('bc,bhwc->bhw', g.astype(np.float64).mean(axis=(1, 2)), A.astype(np.float64)), 0)
    assert rel(cam.cpu().numpy(), want) < 1e-5


def test_argmax_ties_and_max_of_zeros(pkg):
    x = np.zeros((3, 700), np.float32)
    x[0, [5, 600]] = 2.0          # tie: the first index wins (tf.math.argmax)
    x[1, 699] = 1.0
    xd = torch.from_numpy(x).cuda()
    idx = torch.empty(3, dtype=torch.int32, device="cuda")
    pkg._lib.check(pkg._lib.LIB.sky_argmax_rows(xd.data_ptr(), idx.data_ptr(), 3, 700, _st()))
    assert idx.cpu().tolist() == [5, 699, 0]
    out = torch.full((1,), 7.0, device="cuda")
    z = torch.zeros(1000, device="cuda")
    pkg._lib.check(pkg._lib.LIB.sky_max_nonneg(z.data_ptr(), out.data_ptr(), 1000, _st()))
    assert float(out) == 0.0


def test_dense_single_row_and_exact_tile(pkg):
    rng = np.random.default_rng(1)
    for B, K, N in ((1, 64, 256), (3, 4, 512)):
        d = pkg.sunpose_net.Dense(N)
        d.build((B, K))
        W = rng.standard_normal((K, N)).astype(np.float32)
        d.kernel.copy_(torch.from_numpy(W))
        x = rng.standard_normal((B, K)).astype(np.float32)
        dy = rng.standard_normal((B, N)).astype(np.float32)
        assert rel(d(torch.from_numpy(x).cuda()).cpu().numpy(), x.astype(np.float64) @ W) < 1e-5
        assert rel(d.backward_data(torch.from_numpy(dy).cuda()).cpu().numpy(), dy.astype(np.float64) @ W.T) < 1e-5


def test_fused_blend_equals_unfused_tail(pkg):
    """sky_conv2d_fwd_blend (inference) and conv1_u followed by sky_blend_split (train / test step) are the same arithmetic."""
    rng = np.random.default_rng(2)
    B, H, W = 1, 16, 64
    x = rng.standard_normal((B, H, W, 32)).astype(np.float32)
    sky = rng.uniform(0.6, 1.05, (B, H, W, 3)).astype(np.float32)
    rad = rng.uniform(0, 1, (B, H, W, 3)).astype(np.float32)
    conv = pkg.ops.conv2d(output_channels=3, k_h=7, k_w=7, strides=1)
    xd, skyd, radd = (torch.from_numpy(a).cuda() for a in (x, sky, rad))
    fused = conv(xd, leaky_slope=0.1, residual=radd, relu=True, log_decompress=True, blend=(skyd, 0.12))
    sun = conv(xd, leaky_slope=0.1, residual=radd, relu=True)
    yg, yl = torch.empty_like(sun), torch.empty_like(sun)
    pkg._lib.check(pkg._lib.LIB.sky_blend_split(skyd.data_ptr(), sun.data_ptr(), 0.12, yg.data_ptr(), yl.data_ptr(), None, None, None,
                                                B * H * W, _st()))
    assert torch.equal(fused, yl)


@pytest.mark.parametrize("C,F,k,s,shape", [(256, 512, 4, 1, (2, 4, 16)), (128, 256, 4, 2, (2, 8, 32)), (8, 64, 4, 2, (1, 32, 128)), (256, 384, 3, 1, (1, 4, 16))])
def test_wide_and_split_k_convs(pkg, C, F, k, s, shape):
    """sunRadNet / discriminator shapes on the direct kernel: even kernels, stride 2, split-K with the finalize pass, more than 256
    filters as equal slices in one launch (512) or as unequal slices in a loop (384), C % 4 == 0 inputs."""
    rng = np.random.default_rng(C + F)
    B, h, w = shape
    x = rng.standard_normal((B, h, w, C)).astype(np.float32)
    kern = (rng.standard_normal((k, k, C, F)) / np.sqrt(k * k * C)).astype(np.float32)
    bias = rng.standard_normal(F).astype(np.float32)
    layer = pkg.ops.conv2d(output_channels=F, k_h=k, k_w=k, strides=s, kernel_initializer=kern, bias_initializer=bias)
    got = layer(torch.from_numpy(x).cuda(), leaky_slope=0.3).cpu().numpy()
    want = M.leaky_relu(M.conv2d_same(x, kern, bias, stride=s, acc_dtype=torch.float64), 0.3).numpy()
    assert got.shape == want.shape
    assert rel(got, want) < 1.5e-3, rel(got, want)
