"""Host-side mirror of the output step of the reference's utils.py: ``writeHDR(arr, outfilename, imgshape)`` (utils.py:62-84; inference.py:156)
for the ``.hdr`` extension — the Radiance RGBE picture format cv2.imwrite produces.  The shared-exponent encoding runs on the device
(`sky_rgbe_encode`), the host writes the header and flat (not run-length-encoded) scanlines, which every Radiance reader accepts.
``readHDR`` / ``rgbe_encode_numpy`` / ``rgbe_decode_numpy`` are the CPU counterparts used by the tests."""
from __future__ import annotations

import numpy as np
import torch

from ._lib import LIB, check
from .distortion_aware_ops import _require_cuda, _stream


def rgbe_encode_numpy(rgb):
    rgb = np.maximum(np.asarray(rgb, np.float32), 0)
    v = rgb.max(axis=-1)
    m, e = np.frexp(v)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.where(v >= 1e-32, m.astype(np.float32) * np.float32(256.0) / v, 0).astype(np.float32)
    out = np.zeros(rgb.shape[:-1] + (4,), np.uint8)
    out[..., :3] = (rgb * s[..., None]).astype(np.int32).astype(np.uint8)
    out[..., 3] = np.where(v >= 1e-32, e + 128, 0).astype(np.uint8)
    return out


def rgbe_decode_numpy(rgbe):
    rgbe = np.asarray(rgbe, np.uint8)
    f = np.ldexp(1.0, rgbe[..., 3].astype(np.int32) - (128 + 8)).astype(np.float32)
    return np.where(rgbe[..., 3:4] == 0, 0, rgbe[..., :3].astype(np.float32) * f[..., None]).astype(np.float32)


def rgbe_encode(arr, bgr=True):
    """[..., H, W, 3] float32 CUDA tensor -> [..., H, W, 4] uint8 CUDA tensor (R, G, B, E)."""
    arr = _require_cuda(arr, "arr")
    out = torch.empty(arr.shape[:-1] + (4,), dtype=torch.uint8, device=arr.device)
    check(LIB.sky_rgbe_encode(arr.data_ptr(), out.data_ptr(), arr.numel() // 3, int(bool(bgr)), _stream()))
    return out


def writeHDR(arr, outfilename, imgshape=None):
    """utils.writeHDR for '*.hdr' (utils.py:83-84): arr = one BGR panorama [H, W, 3] (numpy or CUDA tensor), as inference.py holds it."""
    if not outfilename.endswith(".hdr"):
        raise NotImplementedError("only the Radiance .hdr branch of utils.writeHDR is mirrored (the .exr branch is commented out there)")
    if isinstance(arr, torch.Tensor) and arr.is_cuda:
        rgbe = rgbe_encode(arr.reshape(arr.shape[-3:]), bgr=True).cpu().numpy()
    else:
        a = np.asarray(arr, np.float32).reshape(np.asarray(arr).shape[-3:])
        rgbe = rgbe_encode_numpy(a[..., ::-1])
    H, W = rgbe.shape[:2]
    with open(outfilename, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n")
        f.write(f"-Y {H} +X {W}\n".encode("ascii"))
        f.write(rgbe.tobytes())


def readHDR(filename):
    """Flat-scanline Radiance reader -> RGB float32 [H, W, 3] (for round-trip tests)."""
    with open(filename, "rb") as f:
        data = f.read()
    head, _, rest = data.partition(b"\n\n")
    if not head.startswith(b"#?RADIANCE"):
        raise ValueError("not a Radiance picture")
    dims, _, pix = rest.partition(b"\n")
    tok = dims.split()
    H, W = int(tok[1]), int(tok[3])
    return rgbe_decode_numpy(np.frombuffer(pix[:H * W * 4], np.uint8).reshape(H, W, 4))
