"""CPU checks of the oracle's restatement of the sun branch and the train / test tail against independent formulations (numpy /
scipy / closed forms), so that the GPU parity tests compare against something that has itself been pinned.  No GPU needed."""
import numpy as np
import torch
from scipy import ndimage

from oracle import da_oracle as O
from oracle import model_oracle as M


def test_apply_rf_is_piecewise_linear_interpolation():
    rng = np.random.default_rng(0)
    k = 1024
    crf = np.sort(rng.uniform(0, 1, (3, k)), axis=1).astype(np.float64)
    x = rng.uniform(0, 1, (3, 50)).astype(np.float64)
    got = M.apply_rf(torch.from_numpy(x), torch.from_numpy(crf)).numpy()
    want = np.stack([np.interp(x[b], np.linspace(0, 1, k), crf[b]) for b in range(3)])
    assert np.abs(got - want).max() < 1e-12
    assert np.allclose(M.apply_rf(torch.ones(3, 1, dtype=torch.float64), torch.from_numpy(crf)).numpy()[:, 0], crf[:, -1])   # x = 1: clamped neighbour


def test_log_codec_round_trip_and_kl():
    x = torch.linspace(0, 3e4, 1000, dtype=torch.float64)
    assert torch.allclose(M.hdr_log_decompression(M.hdr_log_compression(x)), x, rtol=1e-10, atol=1e-9)
    assert abs(float(M.hdr_log_compression(torch.tensor([1.0], dtype=torch.float64))) - 1.0) < 1e-12          # log(11)/log(11)
    p = torch.softmax(torch.randn(4, 100, dtype=torch.float64), -1)
    q = torch.softmax(torch.randn(4, 100, dtype=torch.float64), -1)
    want = torch.nn.functional.kl_div(q.log(), p, reduction="batchmean")
    assert abs(float(M.kl_divergence(p, q)) - float(want)) < 1e-12
    assert abs(float(M.kl_divergence(p, p))) < 1e-15


def test_gaussian_filter_and_dog_against_scipy():
    rng = np.random.default_rng(1)
    img = rng.standard_normal((1, 12, 20, 2))
    for sigma in (1.2489996, 3.0900156):
        g = np.exp(-np.array([1.0, 0.0, 1.0]) / (2 * sigma * sigma))
        g /= g.sum()
        want = np.stack([ndimage.correlate(img[0, :, :, c], np.outer(g, g), mode="mirror") for c in range(2)], -1)[None]   # 'mirror' == tf REFLECT
        got = M.gaussian_filter2d(torch.from_numpy(img), sigma).numpy()
        assert np.abs(got - want).max() < 1e-12
    # DoG is linear: DoG(a) - DoG(b) == DoG(a - b) (the identity the fused L1 kernel relies on to rounding)
    a, b = torch.from_numpy(rng.standard_normal((1, 8, 8, 1))), torch.from_numpy(rng.standard_normal((1, 8, 8, 1)))
    for da, db, dd in zip(M.dog(a), M.dog(b), M.dog(a - b)):
        assert torch.allclose(da - db, dd, atol=1e-12)
    # x2 bilinear resize with half-pixel centres == scipy zoom of order 1 on the half-pixel grid
    up = O.resize_bilinear(a, 16, 16).numpy()[0, :, :, 0]
    assert abs(up[0, 0] - float(a[0, 0, 0, 0])) < 1e-12 and abs(up[1, 1] - float(0.75 * 0.75 * a[0, 0, 0, 0] + 0.75 * 0.25 * (a[0, 0, 1, 0] + a[0, 1, 0, 0]) + 0.0625 * a[0, 1, 1, 0])) < 1e-12


def test_grad_cam_layer_closed_form_and_pool_routing():
    # y_c = sum_c v_c * mean_hw(A_c)  =>  d y_c / dA = v_c / (h w) everywhere, weights = v / (h w), cam = relu(sum_c v_c A_c) / (h w)
    rng = np.random.default_rng(2)
    A = torch.from_numpy(rng.standard_normal((2, 4, 6, 5))).requires_grad_(True)
    v = torch.from_numpy(rng.standard_normal(5))
    y_c = (A.mean(dim=(1, 2)) * v).sum(-1)
    cam = M.grad_cam_layer(y_c, A)
    want = torch.relu((A.detach() * v).sum(-1) / 24.0).unsqueeze(-1)
    assert torch.allclose(cam, want, atol=1e-14)
    # max-pool gradient goes to the FIRST maximum of a window in scan order (TensorFlow MaxPoolGrad; ReLU outputs tie at 0 all the time)
    x = torch.zeros(1, 2, 2, 1, dtype=torch.float64, requires_grad=True)
    M.maxpool2x2_same(x).sum().backward()
    assert x.grad.flatten().tolist() == [1.0, 0.0, 0.0, 0.0]


def test_batch_norm_fold_identity_and_valid_conv_is_cropped_same():
    rng = np.random.default_rng(3)
    x = torch.from_numpy(rng.standard_normal((2, 8, 16, 6)))
    kern = torch.from_numpy(rng.standard_normal((4, 4, 6, 5)) * 0.1)
    g, b, m, v = (torch.from_numpy(a) for a in (1 + 0.1 * rng.standard_normal(5), rng.standard_normal(5), rng.standard_normal(5), rng.uniform(0.5, 2, 5)))
    zero = torch.zeros(5, dtype=torch.float64)
    ref = M.batch_norm_inference(M.conv2d_same(x, kern, zero, stride=2, acc_dtype=torch.float64), g, b, m, v)
    s = g * torch.rsqrt(v + 1e-3)
    folded = M.conv2d_same(x, kern * s, b - m * s, stride=2, acc_dtype=torch.float64)      # what sky_bn_fold feeds the conv kernel
    assert torch.allclose(ref, folded, atol=1e-12)
    # Conv2D(1, 4) VALID == the SAME conv cropped by one pixel in front (what discriminator.model does on the GPU)
    k1 = torch.from_numpy(rng.standard_normal((4, 4, 6, 1)))
    same = M.conv2d_same(x, k1, torch.zeros(1, dtype=torch.float64), acc_dtype=torch.float64)
    valid = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), k1.permute(3, 2, 0, 1)).permute(0, 2, 3, 1)
    assert torch.allclose(same[:, 1:8 - 2, 1:16 - 2, :], valid, atol=1e-12)


def test_generator_inference_oracle_shapes_and_ranges():
    rng = np.random.default_rng(4)
    H, W = 32, 128
    ldr = (np.round(255 * rng.uniform(0, 1, (1, H, W, 3))) / 255).astype(np.float32)
    d = M.generator_inference(ldr, M.random_full_generator_weights(3, H, W), M.random_sunpose_weights(5, H, W), details=True)
    assert tuple(d["y_lin"].shape) == (1, H, W, 3) and bool(torch.isfinite(d["y_lin"]).all()) and float(d["y_lin"].min()) >= 0
    assert abs(float(d["sm"].sum()) - 1) < 1e-5 and float(d["alpha"].min()) >= 0 and float(d["alpha"].max()) <= 1
    assert [tuple(c.shape) for c in d["cams"]] == [(1, H, W, 1), (1, H // 2, W // 2, 1), (1, H // 4, W // 4, 1)]
    assert float(d["sun_rad_gamma"].max()) <= float(M.hdr_log_compression(torch.tensor(30000.0))) + 1e-6          # the 30000 clamp


# ---- golden vectors produced by the reference's OWN tf_utils.py / sunrad_net.py over the TF shim (tests/golden/make_golden_utils.py) ----
def _utils_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "utils_golden.npz"))


def test_oracle_matches_reference_tf_utils_vectors(pkg):
    g = _utils_golden()
    T = torch.from_numpy
    # log codec (tf_utils.py:263-280)
    assert np.allclose(M.hdr_log_compression(T(g["codec_x"])).numpy(), g["codec_compressed"], rtol=2e-6, atol=1e-7)
    assert np.allclose(M.hdr_log_decompression(T(g["codec_compressed"])).numpy(), g["codec_roundtrip"], rtol=2e-5, atol=1e-6)
    # camera response lookup (tf_utils.py:191-255): fp32 oracle against the reference's fp32 ops
    got = M.apply_rf(T(g["rf_x"]), T(g["rf_crf"])).numpy()
    assert np.abs(got - g["rf_y"]).max() < 2e-6
    # sun-position bins / directions (tf_utils.py:95-129) as restated in <package>/dataset.py
    h, w = (int(v) for v in g["bins_hw"])
    assert np.abs(pkg.dataset.sunpose_bins(h, w) - g["bins"]).max() < 1e-6
    for p, want in zip(g["s2w_pts"], g["s2w"]):
        assert np.abs(pkg.dataset.sphere2world((p[0], p[1]), h, w) - want).max() < 1e-6


def test_oracle_matches_reference_sunrad_tail():
    """sunRadNet.call's arithmetic (sunrad_net.py:56-71) executed from the reference source with the conv stack stubbed out: the
    oracle's radiance function, its epsilon placement and the 30000 clamp."""
    g = _utils_golden()
    heads = torch.from_numpy(g["rad_heads"])
    x = torch.from_numpy(g["rad_x"])
    gamma_in = torch.sigmoid(heads[:, 0]).reshape(-1, 1, 1, 1)
    beta_in = torch.sigmoid(heads[:, 1]).reshape(-1, 1, 1, 1)
    assert np.allclose(gamma_in.numpy(), g["rad_gamma_in"], rtol=1e-6) and np.allclose(beta_in.numpy(), g["rad_beta_in"], rtol=1e-6)
    eps = 1e-5
    v = torch.exp(-((1.0 - x) ** 2.0) / (beta_in + eps)) * gamma_in / (beta_in * torch.sqrt(torch.tensor(float(np.float32(np.pi)))) + eps)
    v = torch.where(v > 30000.0, torch.full_like(v, 30000.0), v)
    assert np.allclose(v.numpy(), g["rad_y"], rtol=1e-5, atol=1e-7)
    # and the same through the oracle function the GPU tests use (identity conv stack: weights that produce exactly these head values)
    w = {"gamma": (np.zeros((1, 1), np.float32), np.zeros(1, np.float32)), "beta": (np.zeros((1, 1), np.float32), np.zeros(1, np.float32))}
    for b in range(x.shape[0]):
        wb = dict(w)
        wb["gamma"] = (np.zeros((4 * 16 * 16, 1), np.float32), g["rad_heads"][b, 0:1])
        wb["beta"] = (np.zeros((4 * 16 * 16, 1), np.float32), g["rad_heads"][b, 1:2])
        for name, cin, f in (("d1", 6, 2), ("d2", 2, 2), ("d3", 2, 2), ("d4", 2, 16)):
            wb[name] = {"kernel": np.zeros((4, 4, cin, f), np.float32)}
        plz = torch.zeros(1, 32, 128, 6)
        xb = torch.nn.functional.interpolate(x[b:b + 1].permute(0, 3, 1, 2), size=(32, 128)).permute(0, 2, 3, 1)
        got = M.sunrad_net(xb, plz, wb)
        want = torch.nn.functional.interpolate(torch.from_numpy(g["rad_y"][b:b + 1]).permute(0, 3, 1, 2), size=(32, 128)).permute(0, 2, 3, 1)
        assert np.allclose(got.numpy(), want.numpy(), rtol=1e-5, atol=1e-7)


def test_oracle_matches_reference_train_vectors(pkg):
    """train.vMF and train._preprocessing executed from the reference's own train.py (random draws recorded, JPEG = identity)."""
    g = _utils_golden()
    T = torch.from_numpy
    H, W = 32, 128
    bins = pkg.dataset.sunpose_bins(H, W)
    for p, want in zip(g["vmf_pts"], g["vmf"]):
        got = pkg.dataset.vMF(float(p[0]), float(p[1]), H, W, bins=bins)
        assert np.abs(got - want).max() <= 2e-5 * want.max() and abs(float(got.sum()) - 1) < 1e-5
        assert int(got.argmax()) == int(want.argmax())
    sigma_s = (np.float32(0.08 / 6) * g["pre_u_s"]).reshape(4, 3)          # train.py:67
    sigma_c = (np.float32(0.005) * g["pre_u_c"]).reshape(4, 3)             # train.py:69
    hdr_t, ldr = M.ldr_synth(T(g["pre_hdr"]), T(g["pre_t"]), T(g["pre_crf"]), T(sigma_s), T(sigma_c), T(g["pre_n_s"]), T(g["pre_n_c"]),
                             quantize=True)
    assert np.allclose(hdr_t.numpy(), g["pre_hdr_t"], rtol=2e-6, atol=1e-7)
    code_got, code_want = np.round(ldr.numpy() * 255), np.round(g["pre_ldr"] * 255)
    assert (code_got != code_want).mean() < 1e-3                            # 8-bit codes: identical except ties at x.5 under fp32 reordering
    assert np.abs(code_got - code_want).max() <= 1


def test_oracle_matches_reference_generator_glue():
    """The non-conv glue of generator.model (decoder tails, sun_rad_estimation's normalisation / resize / concat order, blending)
    executed from the reference's own generator.py with recording stand-ins for the layers."""
    g = _utils_golden()
    T = torch.from_numpy
    conv, inp = T(g["gen_conv_out"]), T(g["gen_input"])
    want_tail = torch.relu(inp + M.leaky_relu(conv, 0.1)).numpy()                      # what decode_branch computes after conv1_*
    assert np.allclose(g["gen_sky_decode"], want_tail, atol=1e-7) and np.allclose(g["gen_sun_decode"], want_tail, atol=1e-7)
    assert np.allclose(g["gen_blend"], (conv + inp).numpy(), atol=1e-7)
    # sun_rad_estimation: x = pred / max(pred) over the WHOLE batch; plz = [ldr, cam1, resize(cam2), resize(cam3)]; output tiled x3
    pred = T(g["sre_pred"])
    assert np.allclose(g["sre_x"], (pred / pred.max()).numpy(), rtol=1e-6)
    H, W = g["sre_ldr"].shape[1:3]
    plz = torch.cat([T(g["sre_ldr"]), T(g["sre_cam1"]), M.O.resize_bilinear(T(g["sre_cam2"]), H, W), M.O.resize_bilinear(T(g["sre_cam3"]), H, W)], -1)
    assert np.allclose(g["sre_plz"], plz.numpy(), atol=1e-6)
    assert g["sre_out"].shape[-1] == 3 and np.allclose(g["sre_out"], np.repeat(g["sre_x"] * 2, 3, axis=-1), rtol=1e-6)


def test_oracle_matches_reference_grad_cam_arithmetic():
    """grad_cam.layer executed from the reference's own grad_cam.py with a recorded gradient in place of tf.gradients: channel mean
    over (h, w), einsum over channels, ReLU, trailing axis — the arithmetic of sky_gradcam and of the oracle's grad_cam_layer."""
    g = _utils_golden()
    A, grad = g["cam_A"].astype(np.float64), g["cam_grad"].astype(np.float64)
    want = np.maximum(np.einsum("bc,bhwc->bhw", grad.mean(axis=(1, 2)), A), 0)[..., None]
    assert g["cam_out"].shape == want.shape and np.allclose(g["cam_out"], want, rtol=1e-5, atol=1e-7)
    At = torch.from_numpy(g["cam_A"]).double().requires_grad_(True)
    y_c = (At * torch.from_numpy(g["cam_grad"]).double()).sum(dim=(1, 2, 3))          # d y_c / dA == the recorded gradient
    assert np.allclose(M.grad_cam_layer(y_c, At).detach().numpy(), g["cam_out"], rtol=1e-5, atol=1e-7)


def test_sunpose_wiring_matches_reference_call_order():
    """The reference's own sunpose_net.py run with logging stand-ins for every layer (tests/golden/make_golden_utils.py) against the
    order the oracle (and the package mirror) apply them in."""
    g = _utils_golden()
    order = [str(v) for v in g["sunpose_call_order"]]
    assert order == ["sunlayer1", "pool1_s", "sunlayer2", "pool2_s", "sunlayer3", "pool3_s", "flat", "fc1", "actv1_s", "fc2", "actv2_s",
                     "softmax", "layer.conv1", "layer.norm1", "layer.actv1", "layer.conv2", "layer.norm2", "layer.actv2"]
    assert int(g["sunpose_n_acts"][0]) == 3
