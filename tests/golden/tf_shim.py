"""A numpy + glibc stand-in for the few dozen TensorFlow symbols that /root/reference/distortion_aware_ops.py touches.

TEST INFRASTRUCTURE ONLY (used by make_golden.py in the build container; never imported by the product).

TensorFlow cannot be installed in the build container (no wheel, no network), so the reference's *own source file* is
executed over this stand-in to produce the golden vectors that pin the oracle.  The stand-in follows the eager-mode
TF2 semantics that matter for bit-exactness of the sampling geometry:

* every op is a separately rounded IEEE fp32 numpy op (numpy never contracts to FMA across ops);
* python scalars meeting a Tensor are converted to the Tensor's dtype first (``ops.convert_to_tensor(y, dtype_hint=x.dtype)``);
  bare python floats become fp32, bare python ints int32;
* scalar transcendental ops call glibc's float routines (tanf, cosf, sinf, atan2f, asinf) through ctypes, which is what
  Eigen's scalar path (size-1 eager tensors) resolves to on CPU;
* ``tf.image.resize`` BILINEAR is the TF2 half-pixel-centre kernel (scale=in/out, src=(dst+0.5)*scale-0.5,
  lower=max(floor,0), upper=min(ceil,n-1), lerp=src-floor; top/bottom lerp then vertical lerp), fp32 throughout;
* ``tf.gather_nd`` raises on out-of-range indices like the CPU kernel does.

It is NOT TensorFlow: matmul accumulation order (BLAS) is unspecified in both, which is why HDR values carry a
tolerance while indices/offsets/weights are compared bit for bit.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import math
import sys
import types

import numpy as np

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
for _n in ("tanf", "cosf", "sinf", "asinf"):
    getattr(_libm, _n).restype = ctypes.c_float
    getattr(_libm, _n).argtypes = [ctypes.c_float]
_libm.atan2f.restype = ctypes.c_float
_libm.atan2f.argtypes = [ctypes.c_float, ctypes.c_float]

float32 = np.dtype(np.float32)
int32 = np.dtype(np.int32)

# gather_nd index tensors seen by the shim, in call order (make_golden.py reads these to pin the indices)
GATHER_LOG: list = []
TRACE: dict = {}


class _Shape(tuple):
    def as_list(self):
        return list(self)


def _to_np(v, dtype=None):
    if isinstance(v, Tensor):
        a = v._a
    elif isinstance(v, (list, tuple)):
        a = np.array([_to_np(e) for e in v])
    else:
        a = v
    if dtype is not None:
        return np.asarray(a, dtype=dtype)
    if isinstance(a, np.ndarray) or isinstance(a, np.generic):
        a = np.asarray(a)
        if a.dtype == np.float64:
            return a.astype(np.float32)
        if a.dtype == np.int64:
            return a.astype(np.int32)
        return a
    if isinstance(a, bool):
        return np.asarray(a)
    if isinstance(a, int):
        return np.asarray(a, dtype=np.int32)
    if isinstance(a, float):
        return np.asarray(a, dtype=np.float32)
    return np.asarray(a)


class Tensor:
    __array_priority__ = 1000
    __array_ufunc__ = None      # numpy scalars (np.int32 h, np.float64 theta) defer to our reflected operators

    def __init__(self, a):
        self._a = np.asarray(a)

    # -- structure ---------------------------------------------------------------------------------------------
    @property
    def dtype(self):
        return self._a.dtype

    @property
    def shape(self):
        return _Shape(self._a.shape)

    def get_shape(self):
        return _Shape(int(s) for s in self._a.shape)

    def numpy(self):
        return self._a if self._a.ndim else self._a[()]

    def __getitem__(self, idx):
        return Tensor(self._a[idx])

    def __iter__(self):
        for i in range(self._a.shape[0]):
            yield Tensor(self._a[i])

    def __len__(self):
        return self._a.shape[0]

    def __bool__(self):
        return bool(self._a)

    def __repr__(self):
        return f"shim.Tensor({self._a!r})"

    # -- arithmetic: the other operand adopts this tensor's dtype (TF's dtype_hint rule) -----------------------------
    def _other(self, o):
        if isinstance(o, Tensor):
            if o.dtype != self.dtype:
                raise TypeError(f"dtype mismatch {self.dtype} vs {o.dtype} (TensorFlow would raise too)")
            return o._a
        return np.asarray(_to_np(o), dtype=self.dtype)

    def __add__(self, o): return Tensor(self._a + self._other(o))
    def __radd__(self, o): return Tensor(self._other(o) + self._a)
    def __sub__(self, o): return Tensor(self._a - self._other(o))
    def __rsub__(self, o): return Tensor(self._other(o) - self._a)
    def __mul__(self, o): return Tensor(self._a * self._other(o))
    def __rmul__(self, o): return Tensor(self._other(o) * self._a)
    def __truediv__(self, o): return Tensor(self._a / self._other(o))
    def __rtruediv__(self, o): return Tensor(self._other(o) / self._a)
    def __neg__(self): return Tensor(-self._a)
    def __gt__(self, o): return Tensor(self._a > self._other(o))
    def __ge__(self, o): return Tensor(self._a >= self._other(o))
    def __lt__(self, o): return Tensor(self._a < self._other(o))
    def __le__(self, o): return Tensor(self._a <= self._other(o))


def _t(v, dtype=None):
    return v if (isinstance(v, Tensor) and dtype is None) else Tensor(_to_np(v, dtype))


def _libm1(name):
    f = getattr(_libm, name)

    def op(x):
        a = _to_np(x)
        assert a.dtype == np.float32
        out = np.empty_like(a)
        flat_in, flat_out = a.reshape(-1), out.reshape(-1)
        for i in range(flat_in.size):
            flat_out[i] = f(float(flat_in[i]))
        return Tensor(out)
    return op


def _atan2(y, x):
    ya, xa = _to_np(y), _to_np(x)
    assert ya.dtype == np.float32 and xa.dtype == np.float32
    ya, xa = np.broadcast_arrays(ya, xa)
    out = np.empty(ya.shape, np.float32)
    fo = out.reshape(-1)
    for i, (p, q) in enumerate(zip(ya.reshape(-1), xa.reshape(-1))):
        fo[i] = _libm.atan2f(float(p), float(q))
    return Tensor(out)


def _binary(npop, trace_key=None):
    def op(x, y, name=None):
        x = _t(x)
        r = npop(x._a, x._other(y))
        if trace_key is not None and r.ndim == 4:
            TRACE.setdefault(trace_key, []).append(r.copy())
        return Tensor(r)
    return op


def divide(x, y, name=None):
    # tf.divide: a non-tensor x takes y's dtype if y is a tensor, else the default conversion; then x / y
    if not isinstance(x, Tensor):
        x = _t(x, y.dtype if isinstance(y, Tensor) else None)
    return x / y


def constant(v, dtype=None):
    return _t(v, dtype)


def convert_to_tensor(v, dtype=None):
    return _t(v, dtype)


def cast(x, dtype):
    a = _to_np(x)
    dt = np.dtype(dtype)
    if dt.kind == "i" and a.dtype.kind == "f":
        return Tensor(np.trunc(a).astype(dt))
    return Tensor(a.astype(dt))


def cross(a, b):
    a, b = _to_np(a), _to_np(b)
    o0 = a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1]
    o1 = a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2]
    o2 = a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]
    return Tensor(np.stack([o0, o1, o2], axis=-1).astype(np.float32))


def squeeze(x, axis=None):
    return Tensor(np.squeeze(_to_np(x), axis=axis))


def stack(vals, axis=0):
    return Tensor(np.stack([_to_np(v) for v in vals], axis=axis))


def transpose(x, perm=None):
    return Tensor(np.transpose(_to_np(x), perm))


def expand_dims(x, axis):
    return Tensor(np.expand_dims(_to_np(x), axis))


def reshape(x, shape):
    return Tensor(np.reshape(_to_np(x), [int(s) for s in shape]))


def tile(x, multiples):
    return Tensor(np.tile(_to_np(x), [int(m) for m in multiples]))


def range_(start, limit=None, delta=1):
    if limit is None:
        start, limit = 0, start
    return Tensor(np.arange(int(_to_np(start)), int(_to_np(limit)), delta, dtype=np.int32))


def meshgrid(x, y):
    X, Y = np.meshgrid(_to_np(x), _to_np(y))
    return Tensor(X), Tensor(Y)


def add_n(vals):
    acc = _to_np(vals[0])
    for v in vals[1:]:
        acc = acc + _to_np(v)        # left to right, each add rounded
    return Tensor(acc)


def clip_by_value(x, lo, hi):
    x = _t(x)
    return Tensor(np.maximum(np.minimum(x._a, x._other(hi)), x._other(lo)))


def where(cond, a, b):
    a = _t(a)
    return Tensor(np.where(_to_np(cond), a._a, a._other(b)))


def floor(x):
    return Tensor(np.floor(_to_np(x)))


def pad(x, paddings):
    return Tensor(np.pad(_to_np(x), [(int(a), int(b)) for a, b in paddings]))


def gather_nd(params, indices):
    p, idx = _to_np(params), _to_np(indices)
    GATHER_LOG.append(idx.copy())
    for d in range(idx.shape[-1]):
        if idx[..., d].min() < 0 or idx[..., d].max() >= p.shape[d]:
            raise IndexError(f"gather_nd index out of range on axis {d} (TF-CPU raises InvalidArgumentError)")
    return Tensor(p[tuple(idx[..., d] for d in range(idx.shape[-1]))])


def matmul(a, b):
    return Tensor(np.matmul(_to_np(a), _to_np(b)))


def bias_add(x, b):
    return Tensor(_to_np(x) + _to_np(b))


def extract_patches(images, sizes, strides, rates, padding):
    a = _to_np(images)
    assert padding == "VALID" and list(rates) == [1, 1, 1, 1]
    _, kh, kw, _ = sizes
    _, sh, sw, _ = strides
    n, H, W, c = a.shape
    oh, ow = (H - kh) // sh + 1, (W - kw) // sw + 1
    out = np.empty((n, oh, ow, kh * kw * c), a.dtype)
    for i in range(kh):
        for j in range(kw):
            out[..., (i * kw + j) * c:(i * kw + j + 1) * c] = a[:, i:i + (oh - 1) * sh + 1:sh, j:j + (ow - 1) * sw + 1:sw, :]
    return Tensor(out)


def _resize_axis(n_in, n_out):
    scale = np.float32(n_in) / np.float32(n_out)
    dst = np.arange(n_out, dtype=np.float32)
    src = (dst + np.float32(0.5)) * scale - np.float32(0.5)
    fl = np.floor(src)
    lower = np.maximum(fl.astype(np.int64), 0)
    upper = np.minimum(np.ceil(src).astype(np.int64), n_in - 1)
    lerp = (src - fl).astype(np.float32)
    return lower, upper, lerp


def resize(images, size, method="bilinear"):
    a = _to_np(images).astype(np.float32)
    oh, ow = int(size[0]), int(size[1])
    ylo, yhi, yl = _resize_axis(a.shape[1], oh)
    xlo, xhi, xl = _resize_axis(a.shape[2], ow)
    xl = xl[None, None, :, None]
    yl = yl[None, :, None, None]
    tl, tr = a[:, ylo][:, :, xlo], a[:, ylo][:, :, xhi]
    bl, br = a[:, yhi][:, :, xlo], a[:, yhi][:, :, xhi]
    top = tl + (tr - tl) * xl
    bot = bl + (br - bl) * xl
    out = top + (bot - top) * yl
    TRACE["resized"] = out.copy()
    return Tensor(out.astype(np.float32))


class Layer:
    """keras.layers.Layer: lazy build on first call, add_weight, dtype float32."""

    def __init__(self, *a, **k):
        self.built = False
        self.dtype = "float32"
        self._weights = {}

    def add_weight(self, name, shape, initializer, trainable=True, dtype=None):
        shape = tuple(int(s) for s in shape)
        init = WEIGHT_INIT.get(name)
        if init is None:
            raise RuntimeError("make_golden.py must provide WEIGHT_INIT[%r]" % name)
        w = np.asarray(init(shape), np.float32)
        assert w.shape == shape
        t = Tensor(w)
        self._weights[name] = t
        return t

    def build(self, input_shape):
        self.built = True

    def __call__(self, x, *a, **k):
        x = _t(x)
        if not self.built:
            self.build(x.get_shape())
            self.built = True
        return self.call(x, *a, **k)


WEIGHT_INIT: dict = {}


def install():
    """Register the stand-in as ``tensorflow`` / ``tensorflow.keras.layers`` and restore ``np.math`` (removed in
    numpy 2; the reference uses np.math.pi at distortion_aware_ops.py:200)."""
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.int32 = float32, int32
    tf.Tensor = Tensor
    tf.divide, tf.constant, tf.convert_to_tensor, tf.cast = divide, constant, convert_to_tensor, cast
    tf.multiply = _binary(np.multiply, 'mul4')   # 4-D products in call() are w0..w3 (:103-106)
    tf.add = _binary(np.add)
    tf.subtract = _binary(np.subtract)
    tf.squeeze, tf.stack, tf.transpose, tf.expand_dims, tf.reshape, tf.tile = squeeze, stack, transpose, expand_dims, reshape, tile
    tf.range, tf.meshgrid, tf.add_n, tf.clip_by_value, tf.where, tf.floor = range_, meshgrid, add_n, clip_by_value, where, floor
    tf.pad, tf.gather_nd, tf.matmul = pad, gather_nd, matmul
    tf.math = types.ModuleType("tensorflow.math")
    tf.math.tan, tf.math.cos, tf.math.sin, tf.math.asin = _libm1("tanf"), _libm1("cosf"), _libm1("sinf"), _libm1("asinf")
    tf.math.atan2 = _atan2
    tf.linalg = types.ModuleType("tensorflow.linalg")
    tf.linalg.cross = cross
    tf.nn = types.ModuleType("tensorflow.nn")
    tf.nn.bias_add = bias_add
    tf.image = types.ModuleType("tensorflow.image")
    tf.image.extract_patches, tf.image.resize = extract_patches, resize
    tf.image.ResizeMethod = types.SimpleNamespace(BILINEAR="bilinear")
    keras = types.ModuleType("tensorflow.keras")
    layers = types.ModuleType("tensorflow.keras.layers")
    layers.Layer = Layer
    keras.layers = layers
    tf.keras = keras
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.keras"] = keras
    sys.modules["tensorflow.keras.layers"] = layers
    if not hasattr(np, "math"):
        np.math = math
    return tf
