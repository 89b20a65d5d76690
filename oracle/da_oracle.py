"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's distortion-aware convolution layers.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs may import this module; the product package never does.

Follows, function by function (paths relative to /root/reference):

  offsets()          distortion_aware_ops.py:198-270   (via oracle/da_offsets_oracle.c, glibc float libm)
  sample()           distortion_aware_ops.py:57-106    (via the same C file)
  sample_np()        same lines, restated a second time in numpy fp32 (cross-check of the C restatement)
  conv2d_forward()   distortion_aware_ops.py:50-123    materialised dataflow: pad -> 4x gather -> blend -> matmul -> bias
  resize_bilinear()  tf.image.resize(BILINEAR) as used at distortion_aware_ops.py:322 (TF2 half-pixel centres;
                     algorithm restated from TensorFlow's published resize_bilinear kernel, not vendored in the reference)
  deconv2d_forward() distortion_aware_ops.py:319-395
  *_backward()       torch autograd through the materialised forward == what TF autodiff computes

Parity status: pinned against tests/golden/da_golden.npz, which is produced by executing the reference's own source over
a numpy/glibc TensorFlow stand-in (tests/golden/make_golden.py).  Real TensorFlow cannot run in the build container,
so with respect to TensorFlow's own numerics parity is unpinned.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

from . import tf32_emu as _emu

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c_oracle(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle_da.so")
    src = os.path.join(_HERE, "da_offsets_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_da.so"], stdout=subprocess.DEVNULL)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c_oracle())
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def offsets(h, w, k, dilation=1, skydome=True, via_f64=False):
    """[h, k*k, 2] fp32 (y, x).  Raises like the reference: AssertionError for even k (:188),
    Exception("undefined coordinates") (:252)."""
    assert k % 2 == 1, "kernel_size must be odd number, current kernel size : {}".format(k)
    out = np.zeros((h, k * k, 2), np.float32)
    fn = _lib().oracle_da_offsets_via_f64 if via_f64 else _lib().oracle_da_offsets
    rc = fn(int(h), int(w), int(k), int(dilation), int(bool(skydome)), _p(out))
    if rc == -1:
        raise Exception("undefined coordinates")
    assert rc == 0, rc
    return out


def pad_amounts(n, k, stride=1):
    """_pad_input (:135-143) for one axis -> (before, after)."""
    same_output = (n + stride - 1) // stride
    valid_output = (n - k + stride) // stride
    if same_output == valid_output:
        return 0, 0
    p = k - 1
    return p // 2, p - p // 2


def sample(h, w, k, off):
    """Corner indices (padded frame) and bilinear weights, [h, w, k*k] each, from the C restatement."""
    n = h * w * k * k
    ints = [np.zeros(n, np.int32) for _ in range(4)]
    flts = [np.zeros(n, np.float32) for _ in range(4)]
    off = np.ascontiguousarray(off, np.float32)
    rc = _lib().oracle_da_sample(int(h), int(w), int(k), _p(off), *[_p(a) for a in ints], *[_p(a) for a in flts])
    shp = (h, w, k * k)
    names = ["y0", "y1", "x0", "x1", "w0", "w1", "w2", "w3"]
    d = {nm: a.reshape(shp) for nm, a in zip(names, ints + flts)}
    d["rc"] = rc
    return d


def sample_np(h, w, k, off):
    """Second, independent restatement of :63-106 in numpy fp32 (each numpy op = one rounded fp32 op)."""
    ph = pad_amounts(h, k)
    pw = pad_amounts(w, k)
    in_h, in_w = h + sum(ph), w + sum(pw)
    f32 = np.float32
    i = np.arange(h, dtype=np.int32)[:, None, None, None]
    j = np.arange(w, dtype=np.int32)[None, :, None, None]
    a = np.arange(k, dtype=np.int32)[None, None, :, None]
    b = np.arange(k, dtype=np.int32)[None, None, None, :]
    y = np.broadcast_to(i + a, (h, w, k, k)).reshape(h, w, k * k).astype(f32)
    x = np.broadcast_to(j + b, (h, w, k, k)).reshape(h, w, k * k).astype(f32)
    y = y + off[:, None, :, 0]
    x = x + off[:, None, :, 1]
    y = np.maximum(np.minimum(y, f32(in_h - 1)), f32(0))
    x = np.where(x < 0, x + f32(in_w), x)
    x = np.where(x > f32(in_w - 1), x - f32(in_w), x)
    y0 = np.floor(y).astype(np.int32)
    x0 = np.floor(x).astype(np.int32)
    y1, x1 = y0 + 1, x0 + 1
    y0 = np.clip(y0, 0, in_h - 1)
    y1 = np.clip(y1, 0, in_h - 1)
    x0_w, x1_w = x0, x1
    x0 = np.where(x0 < 0, x0 + in_w, x0)
    x1 = np.where(x1 < 0, x1 + in_w, x1)
    x0 = np.where(x0 > in_w - 1, x0 - in_w, x0)
    x1 = np.where(x1 > in_w - 1, x1 - in_w, x1)
    fy0, fy1, fx0, fx1 = (v.astype(f32) for v in (y0, y1, x0_w, x1_w))
    return dict(y0=y0, y1=y1, x0=x0, x1=x1,
                w0=(fy1 - y) * (fx1 - x), w1=(fy1 - y) * (x - fx0),
                w2=(y - fy0) * (fx1 - x), w3=(y - fy0) * (x - fx0))


def _as_t(a, dtype=torch.float32):
    return a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


def conv2d_forward(x, kernel, bias, k, dilation=1, skydome=True, off=None, acc_dtype=torch.float32, tf32=None):
    """x [B,h,w,C] NHWC, kernel [k*k*C, F], bias [F] -> [B,h,w,F].  Materialises the same intermediates the reference
    does.  acc_dtype=float64 gives the tolerance anchor (same fp32 indices/weights, fp64 blend and contraction)."""
    x, kernel, bias = _as_t(x), _as_t(kernel), _as_t(bias)
    B, h, w, C = x.shape
    if off is None:
        off = offsets(h, w, k, dilation, skydome)
    s = sample(h, w, k, off)
    assert s["rc"] == 0, "gather index out of range (TF-CPU gather_nd would raise)"
    ph, pw = pad_amounts(h, k), pad_amounts(w, k)
    xp = torch.nn.functional.pad(x, (0, 0, pw[0], pw[1], ph[0], ph[1]))        # :125-150
    in_h, in_w = xp.shape[1], xp.shape[2]
    flat = xp.reshape(B, in_h * in_w, C).to(acc_dtype)
    pix = None
    for (yy, xx, ww) in (("y0", "x0", "w0"), ("y0", "x1", "w1"), ("y1", "x0", "w2"), ("y1", "x1", "w3")):
        idx = torch.from_numpy((s[yy].astype(np.int64) * in_w + s[xx]).reshape(-1))
        p = flat.index_select(1, idx)                                           # :170-184 gather_nd
        term = p * torch.from_numpy(s[ww].reshape(-1)).to(acc_dtype)[None, :, None]   # :112
        pix = term if pix is None else pix + term                               # add_n, left to right
    pix = pix.reshape(B, h * w, k * k * C)                                      # :115
    if tf32 is None:
        tf32 = C > 4
    out = _emu.matmul(pix, kernel.to(acc_dtype), tf32=tf32) + bias.to(acc_dtype)   # :117-119
    return out.reshape(B, h, w, -1)


def resize_bilinear(x, oh, ow):
    """TF2 tf.image.resize(..., BILINEAR), half-pixel centres, no antialias; fp32 ops in TF's order."""
    x = _as_t(x)
    B, h, w, C = x.shape

    def axis(n_in, n_out):
        scale = np.float32(n_in) / np.float32(n_out)
        dst = np.arange(n_out, dtype=np.float32)
        src = (dst + np.float32(0.5)) * scale - np.float32(0.5)
        fl = np.floor(src)
        lo = np.maximum(fl.astype(np.int64), 0)
        hi = np.minimum(np.ceil(src).astype(np.int64), n_in - 1)
        return torch.from_numpy(lo), torch.from_numpy(hi), torch.from_numpy((src - fl).astype(np.float32)).to(x.dtype)

    ylo, yhi, yl = axis(h, oh)
    xlo, xhi, xl = axis(w, ow)
    xl = xl[None, None, :, None]
    yl = yl[None, :, None, None]
    top_rows, bot_rows = x.index_select(1, ylo), x.index_select(1, yhi)
    tl, tr = top_rows.index_select(2, xlo), top_rows.index_select(2, xhi)
    bl, br = bot_rows.index_select(2, xlo), bot_rows.index_select(2, xhi)
    top = tl + (tr - tl) * xl
    bot = bl + (br - bl) * xl
    return top + (bot - top) * yl


def deconv2d_forward(x, kernel, bias, out_hw, k, dilation=1, skydome=True, acc_dtype=torch.float32):
    """distortion_aware_ops.py:319-395: bilinear resize to out_hw, then the stride-1 distortion-aware conv whose offset
    table is built for out_hw (:315)."""
    r = resize_bilinear(_as_t(x), int(out_hw[0]), int(out_hw[1]))
    return conv2d_forward(r, kernel, bias, k, dilation, skydome, acc_dtype=acc_dtype)


def conv2d_backward(x, kernel, bias, dy, k, dilation=1, skydome=True, acc_dtype=torch.float32):
    """(dx, dkernel, dbias) by autograd through the materialised forward."""
    x = _as_t(x).clone().requires_grad_(True)
    kernel = _as_t(kernel).clone().requires_grad_(True)
    bias = _as_t(bias).clone().requires_grad_(True)
    y = conv2d_forward(x, kernel, bias, k, dilation, skydome, acc_dtype=acc_dtype)
    y.backward(_as_t(dy).to(y.dtype))
    return x.grad, kernel.grad, bias.grad


def deconv2d_backward(x, kernel, bias, dy, out_hw, k, dilation=1, skydome=True, acc_dtype=torch.float32):
    x = _as_t(x).clone().requires_grad_(True)
    kernel = _as_t(kernel).clone().requires_grad_(True)
    bias = _as_t(bias).clone().requires_grad_(True)
    y = deconv2d_forward(x, kernel, bias, out_hw, k, dilation, skydome, acc_dtype=acc_dtype)
    y.backward(_as_t(dy).to(y.dtype))
    return x.grad, kernel.grad, bias.grad
