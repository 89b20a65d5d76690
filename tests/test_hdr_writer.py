"""Radiance .hdr output step (utils.writeHDR, inference.py:156): CPU round trip of the RGBE codec and the file writer; the device
encoder is compared bit for bit with the numpy one in the GPU test."""
import numpy as np
import pytest
import torch


def test_rgbe_round_trip_and_file(pkg, tmp_path):
    rng = np.random.default_rng(0)
    img = (rng.uniform(0, 1, (32, 128, 3)) ** 6 * 3e4).astype(np.float32)           # sky + sun dynamic range
    img[0, 0] = 0
    U = pkg.utils
    dec = U.rgbe_decode_numpy(U.rgbe_encode_numpy(img))
    v = img.max(-1, keepdims=True)
    assert np.all(np.abs(dec - img) <= v / 128 + 1e-30)                              # 8-bit mantissa shared per pixel (truncation)
    assert np.array_equal(dec[0, 0], [0, 0, 0])
    path = str(tmp_path / "pred.hdr")
    U.writeHDR(img[..., ::-1], path, (32, 128, 3))                                   # inference.py holds BGR
    back = U.readHDR(path)
    assert back.shape == (32, 128, 3) and np.array_equal(back, dec)
    assert open(path, "rb").read(10) == b"#?RADIANCE"


@pytest.mark.gpu
def test_device_rgbe_encoder_matches_numpy(pkg):
    rng = np.random.default_rng(1)
    img = (rng.uniform(0, 1, (2, 32, 128, 3)) ** 6 * 3e4).astype(np.float32)
    img[0, 0, 0] = 0
    got = pkg.utils.rgbe_encode(torch.from_numpy(img).cuda(), bgr=False).cpu().numpy()
    assert np.array_equal(got, pkg.utils.rgbe_encode_numpy(img))
    got_bgr = pkg.utils.rgbe_encode(torch.from_numpy(np.ascontiguousarray(img[..., ::-1])).cuda(), bgr=True).cpu().numpy()
    assert np.array_equal(got_bgr, got)
