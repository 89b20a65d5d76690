"""GPU parity of the sun-position network forward (sunpose_net.py:54-72) against the oracle, with the distortion-aware
wiring of sunpose_net.py:11,16 and with the plain wiring that is live in the committed reference.
Tolerances (relative L2 vs the fp64 oracle): max-pool exact; Dense (fp32 split-K) 1e-5; activation maps 5e-3 (TF32);
the softmax output 2e-2 in the TF32 mode (4096-way softmax of O(1) logits after two 4096-wide dense layers)."""
import numpy as np
import pytest
import torch

from oracle import model_oracle as M

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def test_maxpool_dense_softmax(pkg):
    sp = pkg.sunpose_net
    rng = np.random.default_rng(0)
    for shape in ((2, 8, 32, 64), (1, 5, 7, 32)):                       # even and odd maps (SAME: ceil)
        x = rng.standard_normal(shape).astype(np.float32)
        want = M.maxpool2x2_same(torch.from_numpy(x)).numpy()
        got = sp.maxpool2d(torch.from_numpy(x).cuda()).cpu().numpy()
        assert np.array_equal(got, want)
    B, K, N = 32, 1000, 777
    x = rng.standard_normal((B, K)).astype(np.float32)
    W = rng.standard_normal((K, N)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    d = sp.Dense(N)
    d.build((B, K))
    d.kernel.copy_(torch.from_numpy(W))
    d.bias.copy_(torch.from_numpy(b))
    want = np.maximum(x.astype(np.float64) @ W.astype(np.float64) + b, 0)
    assert rel_l2(d(torch.from_numpy(x).cuda(), relu=True).cpu().numpy(), want) < 1e-5
    want = x.astype(np.float64) @ W.astype(np.float64) + b
    got = d(torch.from_numpy(x).cuda())
    assert rel_l2(got.cpu().numpy(), want) < 1e-5
    sm = sp.softmax(got).cpu().numpy()
    assert rel_l2(sm, torch.softmax(torch.from_numpy(want), -1).numpy()) < 1e-5


@pytest.mark.parametrize("da", [True, False])
def test_sunpose_forward_vs_oracle(pkg, da):
    rng = np.random.default_rng(1)
    B, H, W = 2, 32, 128
    ldr = (np.round(255 * rng.uniform(0, 1, (B, H, W, 3))) / 255).astype(np.float32)
    w = M.random_sunpose_weights(seed=5, H=H, W=W)
    net = pkg.sunpose_net.model(im_height=H, im_width=W, distortion_aware=da)
    x = torch.from_numpy(ldr).cuda()
    # build + load weights (the reference builds lazily on the first call)
    net.sunposeEstimation(x)
    for name in ("sunlayer1", "sunlayer2", "sunlayer3"):
        layer, d = getattr(net, name), w[name]
        for i, (conv, norm) in enumerate(((layer.conv1, layer.norm1), (layer.conv2, layer.norm2)), start=1):
            if da:
                conv.kernel.copy_(torch.from_numpy(d[f"conv{i}_kernel"]))
                conv.bias.copy_(torch.from_numpy(d[f"conv{i}_bias"]))
            else:
                conv.w.copy_(torch.from_numpy(d[f"conv{i}_kernel"]).reshape(conv.w.shape))
                conv.biases.copy_(torch.from_numpy(d[f"conv{i}_bias"]))
            norm.gamma.copy_(torch.from_numpy(d[f"norm{i}_gamma"]))
            norm.beta.copy_(torch.from_numpy(d[f"norm{i}_beta"]))
    for name in ("fc1", "fc2"):
        getattr(net, name).kernel.copy_(torch.from_numpy(w[name][0]))
        getattr(net, name).bias.copy_(torch.from_numpy(w[name][1]))
    sm, acts = net.sunposeEstimation(x)
    want_sm, want_acts = M.sunpose_estimation(ldr, w, distortion_aware=da, acc_dtype=torch.float64)
    assert tuple(sm.shape) == (B, H * W)
    assert [tuple(a.shape) for a in acts] == [(B, H, W, 32), (B, H // 2, W // 2, 64), (B, H // 4, W // 4, 128)]
    for a, wa in zip(acts, want_acts):
        assert rel_l2(a.cpu().numpy(), wa.numpy()) <= 5e-3, rel_l2(a.cpu().numpy(), wa.numpy())
    assert abs(sm.sum(-1).cpu().numpy() - 1).max() < 1e-4
    assert rel_l2(sm.cpu().numpy(), want_sm.numpy()) <= 2e-2, rel_l2(sm.cpu().numpy(), want_sm.numpy())


@pytest.mark.parametrize("B,K,N", [(32, 1000, 1028), (5, 96, 256), (40, 4096, 512)])
def test_dense_streaming_kernel(pkg, B, K, N):
    """Wide layers (N % 4 == 0, N >= 256) take the bulk-copy weight-streaming kernel: ragged K tail, ragged column tile, B < 32 and
    B > 32 (two row chunks).  fp32 FMA, split-K with atomics: 1e-5 relative L2 against fp64."""
    sp = pkg.sunpose_net
    rng = np.random.default_rng(B + K + N)
    x = rng.standard_normal((B, K)).astype(np.float32)
    W = rng.standard_normal((K, N)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    d = sp.Dense(N)
    d.build((B, K))
    d.kernel.copy_(torch.from_numpy(W))
    d.bias.copy_(torch.from_numpy(b))
    want = x.astype(np.float64) @ W.astype(np.float64) + b
    assert rel_l2(d(torch.from_numpy(x).cuda()).cpu().numpy(), want) < 1e-5
    assert rel_l2(d(torch.from_numpy(x).cuda(), relu=True).cpu().numpy(), np.maximum(want, 0)) < 1e-5


@pytest.mark.parametrize("B,K,N", [(32, 1000, 1028), (5, 256, 96), (40, 4096, 512), (32, 8192, 4096)])
def test_dense_data_gradient_without_transposed_copy(pkg, B, K, N):
    """dx = dy . W^T streamed from W in its own [K, N] layout (128-byte-swizzled TMA tiles, a thread owns two k-rows) against fp64 and
    against the round-1 path (transposed copy + the forward kernel): ragged K / N tails, B < 32 and B > 32, the ReLU mask."""
    import unittest.mock as mock
    sp = pkg.sunpose_net
    rng = np.random.default_rng(B + K + N)
    dy = rng.standard_normal((B, N)).astype(np.float32)
    W = rng.standard_normal((K, N)).astype(np.float32)
    act = rng.standard_normal((B, K)).astype(np.float32)
    d = sp.Dense(N)
    d.build((B, K))
    d.kernel.copy_(torch.from_numpy(W))
    want = dy.astype(np.float64) @ W.astype(np.float64).T
    assert sp.DENSE_BWD_KERNEL == "nt"
    got = d.backward_data(torch.from_numpy(dy).cuda())
    assert rel_l2(got.cpu().numpy(), want) < 1e-5, rel_l2(got.cpu().numpy(), want)
    got_m = d.backward_data(torch.from_numpy(dy).cuda(), act=torch.from_numpy(act).cuda())
    assert rel_l2(got_m.cpu().numpy(), want * (act > 0)) < 1e-5
    with mock.patch.object(sp, "DENSE_BWD_KERNEL", "t"):
        old = d.backward_data(torch.from_numpy(dy).cuda())
    assert rel_l2(got.cpu().numpy(), old.cpu().numpy()) < 1e-5
