"""Host-side mirror of the pieces of the reference's tf_utils.py that sit on the hot path of the train / test step:
``hdr_logCompression`` / ``hdr_logDecompression`` (tf_utils.py:263-280), ``apply_rf`` (:243-255, with interp_1d / sample_1d :191-241),
``DoG`` (:61-73) — the latter as the fused L1-of-differences that train.py:311-319 and train_sun.py:247-253 compute from it."""
from __future__ import annotations

import torch

from ._lib import LIB, check
from .distortion_aware_ops import _ptr, _require_cuda, _stream


def _codec(x, decompress):
    x = _require_cuda(x, "x")
    y = torch.empty_like(x)
    check(LIB.sky_hdr_log_codec(x.data_ptr(), y.data_ptr(), x.numel(), int(decompress), _stream()))
    return y


def hdr_logCompression(x, validDR=10.):
    if validDR != 10.:
        raise NotImplementedError("the reference only uses validDR = 10")
    return _codec(x, False)


def hdr_logDecompression(x, validDR=10.):
    if validDR != 10.:
        raise NotImplementedError("the reference only uses validDR = 10")
    return _codec(x, True)


def ldr_synth(hdr, t, crf, sigma_s=None, sigma_c=None, noise_s=None, noise_c=None, quantize=True):
    """train._preprocessing (train.py:54-94) -> [hdr_t, ldr]; the random draws are inputs (None: that noise term is skipped)."""
    hdr = _require_cuda(hdr, "hdr")
    B, H, W, C = hdr.shape
    t, crf = _require_cuda(t, "t"), _require_cuda(crf, "crf")
    hdr_t, ldr = torch.empty_like(hdr), torch.empty_like(hdr)
    check(LIB.sky_ldr_synth(hdr.data_ptr(), t.data_ptr(), crf.data_ptr(), _ptr(sigma_s), _ptr(sigma_c), _ptr(noise_s), _ptr(noise_c),
                            hdr_t.data_ptr(), ldr.data_ptr(), B, H * W, C, crf.shape[1], int(bool(quantize)), _stream()))
    return [hdr_t, ldr]


def apply_rf(x, rf):
    """tf_utils.apply_rf (:243-255): x [b, ...] in [0, 1], rf [b, k] -> lerp of each sample's response curve at (k-1) x."""
    x = _require_cuda(x, "x")
    b = x.shape[0]
    n = x.numel() // b
    ones = torch.ones(b, dtype=torch.float32, device=x.device)
    xr = x.reshape(b, n, 1, 1)
    return ldr_synth(xr, ones, rf, quantize=False)[1].reshape(x.shape)


class _Accumulator:
    """A few fp64 device scalars the reduction kernels add into (zeroed per use)."""

    def __init__(self, n, device):
        self.buf = torch.zeros(n, dtype=torch.float64, device=device)

    def zero(self):
        self.buf.zero_()
        return self.buf


def reduce_mean_abs_diff(a, b, acc):
    """tf.reduce_mean(tf.abs(a - b)) -> fp64 device scalar view (acc: a zeroed fp64 element)."""
    a, b = _require_cuda(a, "a"), _require_cuda(b, "b")
    check(LIB.sky_loss_reduce(0, a.data_ptr(), b.data_ptr(), a.numel(), acc.data_ptr(), _stream()))
    return acc / a.numel()


def DoG_l1(a, b, acc4):
    """sum_l reduce_mean(|DoG_l(a) - DoG_l(b)|) (train.py:311-319; tf_utils.DoG :61-73) -> fp64 device scalar.  acc4: 4 zeroed fp64."""
    a, b = _require_cuda(a, "a"), _require_cuda(b, "b")
    B, h, w, C = a.shape
    ba = torch.empty((B, 2 * h, 2 * w, C), dtype=torch.float32, device=a.device)
    bb = torch.empty_like(ba)
    check(LIB.sky_dog_base(a.data_ptr(), ba.data_ptr(), B, h, w, C, _stream()))
    check(LIB.sky_dog_base(b.data_ptr(), bb.data_ptr(), B, h, w, C, _stream()))
    check(LIB.sky_dog_l1(ba.data_ptr(), bb.data_ptr(), B, 2 * h, 2 * w, C, acc4.data_ptr(), _stream()))
    return acc4.sum() / ba.numel()


def kl_divergence(y_true, y_pred, acc):
    """tf.keras.losses.KLDivergence()(y_true, y_pred): mean over the batch of sum y_t log(y_t / y_p)."""
    y_true, y_pred = _require_cuda(y_true, "y_true"), _require_cuda(y_pred, "y_pred")
    check(LIB.sky_kl_divergence(y_true.data_ptr(), y_pred.data_ptr(), y_true.numel(), acc.data_ptr(), _stream()))
    return acc / y_true.shape[0]
