"""GPU parity of the kernels around the generator in the train / test step (train.py) against the oracle.
Tolerances (relative, against the fp64 oracle): LDR synthesis / DoRF lookup: fp32 elementwise, 2e-6 absolute on [0, 1] values before
quantisation and exact 8-bit codes after it except where the pre-rounding value sits within 1e-5 of a half-integer;
log codec 1e-6; KL / L1 / LSGAN reductions 1e-5; DoG L1 1e-4; VGG16 pools 5e-3 (TF32, 7 convs); discriminator 5e-3;
total generator loss of the test step 2e-2 (dominated by 1000 x DoG of a TF32 prediction with radiances up to 3e4)."""
import numpy as np
import pytest
import torch

from oracle import model_oracle as M

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def test_ldr_synth_and_apply_rf(pkg):
    rng = np.random.default_rng(0)
    B, H, W, K = 4, 16, 32, 1024
    hdr = (rng.uniform(0, 1, (B, H, W, 3)) ** 3 * 4).astype(np.float32)
    t = (2 ** rng.uniform(-3, 3, B)).astype(np.float32)
    gam = rng.uniform(1.5, 3, (B, 1))
    crf = (np.linspace(0, 1, K)[None, :] ** (1 / gam)).astype(np.float32)
    ss = (0.08 / 6 * rng.uniform(0, 1, (B, 3))).astype(np.float32)
    sc = (0.005 * rng.uniform(0, 1, (B, 3))).astype(np.float32)
    ns, nc = (rng.standard_normal((B, H, W, 3)).astype(np.float32) for _ in range(2))
    T = lambda a: torch.from_numpy(a)
    want_h, want_l = M.ldr_synth(*(T(a).double() for a in (hdr, t, crf, ss, sc, ns, nc)), quantize=False)
    got_h, got_l = pkg.tf_utils.ldr_synth(*(T(a).cuda() for a in (hdr, t, crf, ss, sc, ns, nc)), quantize=False)
    assert np.abs(got_h.cpu().numpy() - want_h.numpy()).max() < 1e-5 * max(1.0, float(want_h.max()))
    assert np.abs(got_l.cpu().numpy() - want_l.numpy()).max() < 2e-6
    got_q = pkg.tf_utils.ldr_synth(*(T(a).cuda() for a in (hdr, t, crf, ss, sc, ns, nc)), quantize=True)[1].cpu().numpy()
    code = want_l.numpy() * 255.0
    safe = np.abs(code - np.floor(code) - 0.5) > 1e-3
    assert np.array_equal(np.round(got_q * 255)[safe], np.round(code)[safe])
    x = rng.uniform(0, 1, (B, 7, 5)).astype(np.float32)
    assert np.abs(pkg.tf_utils.apply_rf(T(x).cuda(), T(crf).cuda()).cpu().numpy() - M.apply_rf(T(x).double(), T(crf).double()).numpy()).max() < 2e-6


def test_codec_and_reductions(pkg):
    rng = np.random.default_rng(1)
    x = (rng.uniform(0, 1, (3, 8, 16, 3)) ** 4 * 50).astype(np.float32)
    xd = torch.from_numpy(x).cuda()
    g = pkg.tf_utils.hdr_logCompression(xd)
    assert rel(g.cpu().numpy(), M.hdr_log_compression(torch.from_numpy(x).double()).numpy()) < 1e-6
    assert rel(pkg.tf_utils.hdr_logDecompression(g).cpu().numpy(), x) < 1e-5
    acc = torch.zeros(8, dtype=torch.float64, device="cuda")
    y = rng.standard_normal(x.shape).astype(np.float32)
    assert abs(float(pkg.tf_utils.reduce_mean_abs_diff(xd, torch.from_numpy(y).cuda(), acc[0:1])) / np.abs(x.astype(np.float64) - y).mean() - 1) < 1e-5
    p = torch.softmax(torch.from_numpy(rng.standard_normal((4, 4096)).astype(np.float32) * 3, ), -1)
    q = torch.softmax(torch.from_numpy(rng.standard_normal((4, 4096)).astype(np.float32)), -1)
    got = float(pkg.tf_utils.kl_divergence(p.cuda(), q.cuda(), acc[1:2]))
    assert abs(got / float(M.kl_divergence(p.double(), q.double())) - 1) < 1e-5


@pytest.mark.parametrize("C", [3, 1])
def test_dog_l1(pkg, C):
    rng = np.random.default_rng(2)
    a = (rng.uniform(0, 1, (2, 16, 32, C)) ** 3 * 30).astype(np.float32)
    b = (a + rng.standard_normal(a.shape) * 0.3).astype(np.float32)
    acc = torch.zeros(4, dtype=torch.float64, device="cuda")
    got = float(pkg.tf_utils.DoG_l1(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), acc))
    want = float(M.dog_l1(torch.from_numpy(a).double(), torch.from_numpy(b).double()))
    assert abs(got / want - 1) < 1e-4, (got, want)


def test_vgg16_and_discriminator(pkg):
    rng = np.random.default_rng(3)
    B, H, W = 2, 32, 128
    img = rng.uniform(0, 1, (B, H, W, 3)).astype(np.float32)
    dd = pkg.vgg16.random_data_dict(1)
    vgg = pkg.vgg16.Vgg16(data_dict=dd)
    got = vgg(torch.from_numpy(img).cuda())
    want = M.vgg16_pools(torch.from_numpy(img), dd, acc_dtype=torch.float64)
    assert [tuple(g.shape) for g in got] == [(B, 16, 64, 64), (B, 8, 32, 128), (B, 4, 16, 256)]
    for g, w in zip(got, want):
        assert rel(g.cpu().numpy(), w.numpy()) <= 5e-3, rel(g.cpu().numpy(), w.numpy())
    wd = M.random_discriminator_weights(4)
    dis = pkg.discriminator.model()
    dis.build(B, H, W)
    dis.set_weights(wd)
    hdr = (rng.uniform(0, 1, (B, H, W, 3)) ** 3 * 5).astype(np.float32)
    got_d = dis([torch.from_numpy(img).cuda(), torch.from_numpy(hdr).cuda()], training=False)
    want_d = M.discriminator(torch.from_numpy(img), torch.from_numpy(hdr), wd, torch.float64)
    assert tuple(got_d.shape) == tuple(want_d.shape) == (B, 1, 13, 1)
    assert rel(got_d.cpu().numpy(), want_d.numpy()) <= 5e-3, rel(got_d.cpu().numpy(), want_d.numpy())


def test_generator_test_step_losses(pkg):
    rng = np.random.default_rng(5)
    B, H, W = 2, 32, 128
    ldr = (np.round(255 * rng.uniform(0, 1, (B, H, W, 3))) / 255).astype(np.float32)
    hdr_t = (rng.uniform(0, 1, (B, H, W, 3)) ** 3 * 3).astype(np.float32)
    gt = torch.softmax(torch.from_numpy(rng.standard_normal((B, H * W)).astype(np.float32) * 4), -1).numpy()
    wg, ws, wd = M.random_full_generator_weights(3, H, W), M.random_sunpose_weights(5, H, W), M.random_discriminator_weights(4)
    dd = pkg.vgg16.random_data_dict(1)
    step = pkg.train.Step(batch_size=B, im_height=H, im_width=W, vgg_data_dict=dd)
    x = torch.from_numpy(ldr).cuda()
    step._sun.sunposeEstimation(x)
    step._gen.set_weights(wg)
    step._sun.set_weights(ws)
    step._dis.set_weights(wd)
    gen_pred, disc_loss = step.test_step([torch.from_numpy(hdr_t).cuda(), x], torch.from_numpy(gt).cuda())
    want = M.generator_test_step(ldr, hdr_t, gt, wg, ws, wd, dd, acc_dtype=torch.float64)
    logl = lambda y: np.log1p(10 * np.asarray(y, np.float64)) / np.log(11.0)
    assert rel(logl(gen_pred[-1].cpu().numpy()), logl(want["y_final_lin"].numpy())) <= 1e-2
    assert rel(logl(gen_pred[2].cpu().numpy()), logl(want["sky_pred_lin"].numpy())) <= 1e-2
    assert rel(logl(gen_pred[3].cpu().numpy()), logl(want["sun_pred_lin"].numpy())) <= 1e-2
    got = {k: float(v) for k, v in step.last_losses.items()}
    for name, tol in (("kl", 2e-2), ("perceptual", 2e-2), ("dog", 2e-2), ("l1", 2e-2), ("gen", 2e-2), ("total", 2e-2)):
        assert abs(got[name] / float(want[name]) - 1) <= tol, (name, got[name], float(want[name]))
    assert abs(float(disc_loss) / float(want["disc"]) - 1) <= 2e-2
