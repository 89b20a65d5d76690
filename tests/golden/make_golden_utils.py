"""Generate tests/golden/utils_golden.npz by executing the reference's own tf_utils.py (hdr_logCompression / hdr_logDecompression,
apply_rf / interp_1d / sample_1d, sphere2world, sunpose_init) and the arithmetic tail of sunrad_net.sunRadNet.call, read
unmodified from /root/reference, over the numpy TensorFlow stand-in of tf_shim.py (extended here with the few extra symbols those
functions touch).  Run in the build container only:   python tests/golden/make_golden_utils.py
The committed .npz pins oracle/model_oracle.py, <package>/dataset.py and the GPU kernels' CPU references (tests/test_oracle_model.py)."""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_shim  # noqa: E402
from tf_shim import Tensor, _to_np  # noqa: E402


def _f32(fn):
    def op(x, *a, **k):
        return Tensor(fn(_to_np(x).astype(np.float32)).astype(np.float32))
    return op


def install():
    tf = tf_shim.install()
    tf.math.log, tf.math.exp = _f32(np.log), _f32(np.exp)
    tf.math.multiply, tf.math.divide, tf.math.subtract = tf.multiply, tf.divide, tf.subtract
    tf.sqrt = _f32(np.sqrt)
    tf.pow = lambda x, y: Tensor(np.power(_to_np(x).astype(np.float32), np.float32(_to_np(y))).astype(np.float32))
    tf.nn.sigmoid = _f32(lambda a: 1.0 / (1.0 + np.exp(-a)))
    tf.shape = lambda x: list(_to_np(x).shape)
    tf.scalar_mul = lambda s, x: Tensor(np.float32(s) * _to_np(x))
    _where3 = tf.where
    tf.where = lambda c, a, b: Tensor(np.where(_to_np(c), np.float32(_to_np(a)) if np.isscalar(a) else _to_np(a), _to_np(b)).astype(np.float32))
    tf.function = lambda f=None, **k: (f if f is not None else (lambda g: g))
    _range = tf.range
    tf.range = lambda *a, dtype=None, **k: (Tensor(_to_np(_range(*a, **k)).astype(np.int32 if dtype is tf.int32 else np.float32))
                                            if dtype is not None else _range(*a, **k))

    def gather_nd(params, indices):
        p, idx = _to_np(params), _to_np(indices).astype(np.int64)
        return Tensor(p[tuple(idx[..., i] for i in range(idx.shape[-1]))])
    tf.gather_nd = gather_nd
    keras = sys.modules["tensorflow.keras"]
    keras.Model = type("Model", (), {"__init__": lambda self, *a, **k: None})
    keras.layers.Conv2D = keras.layers.BatchNormalization = keras.layers.LeakyReLU = keras.layers.Flatten = keras.layers.Dense = (
        lambda *a, **k: None)
    tf.random_normal_initializer = lambda *a, **k: None
    for name in ("tensorflow_addons", "utils", "ops", "cv2", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    return tf


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def train_vectors():
    """train.vMF (train.py:42-52) and train._preprocessing (train.py:54-94) executed from the reference's own train.py.  Its sibling
    model modules are stubbed (they are not touched by these two functions); tf_utils is the real one.  The random draws of
    _preprocessing are replaced by recorded arrays and the JPEG round trip (a CPU codec) by the identity, as SURVEY 8(d) states."""
    tf = sys.modules["tensorflow"]
    tf.config = types.SimpleNamespace(experimental=types.SimpleNamespace(list_physical_devices=lambda kind: [], set_visible_devices=None))
    tf.data = types.SimpleNamespace(AUTOTUNE=-1)
    tf.uint8 = np.uint8
    tf.einsum = lambda eq, a, b: Tensor(np.einsum(eq.replace(" ", ""), _to_np(a), _to_np(b)).astype(np.float32))
    tf.reduce_sum = lambda x, *a, **k: Tensor(np.float32(_to_np(x).sum(dtype=np.float32)))
    tf.reduce_mean = lambda x, *a, **k: Tensor(np.float32(_to_np(x).mean(dtype=np.float32)))
    tf.nn.relu = lambda x: Tensor(np.maximum(_to_np(x), np.float32(0)))
    tf.round = lambda x: Tensor(np.rint(_to_np(x)))
    _cast = tf.cast
    tf.cast = lambda x, dt: (Tensor(_to_np(x).astype(np.uint8)) if dt is np.uint8 else
                             (Tensor(_to_np(x).astype(np.float32)) if _to_np(x).dtype == np.uint8 else _cast(x, dt)))
    tf.image.adjust_jpeg_quality = lambda img, q: img
    draws = {}
    tf.random = types.SimpleNamespace(
        uniform=lambda shape, minval=0.0, maxval=1.0, dtype=None, seed=None: Tensor(draws["uniform"].pop(0)),
        normal=lambda shape, seed=None: Tensor(draws["normal"].pop(0)))
    for name in ("generator", "discriminator", "sunpose_net", "grad_cam", "vgg16"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["vgg16"].Vgg16 = object
    sys.modules["tf_utils"] = load("/root/reference/tf_utils.py", "tf_utils")
    T = load("/root/reference/train.py", "ref_train")
    out = {}
    H, W = T.IMSHAPE[0], T.IMSHAPE[1]
    pts = np.array([[T.AZIMUTH_gt, 8.0], [T.AZIMUTH_gt, 20.5], [10.0, 3.0]], np.float32)
    out["vmf_pts"] = pts
    out["vmf"] = np.stack([T.vMF(float(p[0]), float(p[1]), H, W).numpy() for p in pts])
    rng = np.random.default_rng(7)
    b, K = 4, 1024
    hdr = (rng.uniform(0, 1, (b, 8, 16, 3)) ** 3 * 4).astype(np.float32)
    crf_src = (np.linspace(0, 1, K)[None, :] ** (1 / rng.uniform(1.5, 3, (6, 1)))).astype(np.float32)
    t_src = (2 ** np.linspace(-3, 3, 9)).astype(np.float32)
    picks = iter([3, 0, 5, 1, 8, 2, 2, 7])                       # crf rows for the 4 samples, then exposures
    T.randint = lambda lo, hi: next(picks)
    u_s, u_c = (rng.uniform(0, 1, (b, 1, 1, 3)).astype(np.float32) for _ in range(2))
    n_s, n_c = (rng.standard_normal(hdr.shape).astype(np.float32) for _ in range(2))
    draws["uniform"], draws["normal"] = [u_s, u_c], [n_s, n_c]
    hdr_t, ldr = T._preprocessing(Tensor(hdr), crf_src, t_src)
    out.update(pre_hdr=hdr, pre_crf=crf_src[[3, 0, 5, 1]], pre_t=t_src[[8, 2, 2, 7]], pre_u_s=u_s, pre_u_c=u_c, pre_n_s=n_s, pre_n_c=n_c,
               pre_hdr_t=hdr_t.numpy(), pre_ldr=ldr.numpy())
    return out


def generator_vectors():
    """The non-conv glue of generator.model executed from the reference's own generator.py with every layer replaced by a recording
    stand-in: sky_decode / sun_decode tails (generator.py:120-124, 150-155), sun_rad_estimation's normalisation, resize and concat
    order (:160-167) and blending (:171-175)."""
    tf = sys.modules["tensorflow"]
    tf.nn.leaky_relu = lambda x, a: Tensor(np.where(_to_np(x) > 0, _to_np(x), np.float32(a) * _to_np(x)).astype(np.float32))
    tf.reduce_max = lambda x, *a, **k: Tensor(np.float32(_to_np(x).max()))
    tf.concat = lambda vals, axis=-1: Tensor(np.concatenate([_to_np(v) for v in vals], axis=axis))
    for name in ("ops", "distortion_aware_ops", "sunrad_net"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["sunrad_net"].sunRadNet = object
    sys.modules["tensorflow_addons"].layers = types.SimpleNamespace(InstanceNormalization=lambda *a, **k: None)
    G = load("/root/reference/generator.py", "ref_generator")
    rng = np.random.default_rng(11)
    B, H, W = 2, 8, 16
    m = object.__new__(G.model)
    m.im_height, m.im_width = H, W
    conv_out = rng.standard_normal((B, H, W, 3)).astype(np.float32)
    ident = lambda t, *a, **k: t
    for name in ("conv3_f", "norm3_f", "conv2_f", "norm2_f", "conv3_u", "norm3_u", "conv2_u", "norm2_u"):
        setattr(m, name, ident)
    m.conv1_f = m.conv1_u = lambda t: Tensor(conv_out)
    inp = rng.uniform(0, 1, (B, H, W, 3)).astype(np.float32)
    out = {"gen_conv_out": conv_out, "gen_input": inp}
    out["gen_sky_decode"] = G.model.sky_decode(m, Tensor(np.zeros((B, 2, 4, 8), np.float32)), Tensor(inp)).numpy()
    out["gen_sun_decode"] = G.model.sun_decode(m, Tensor(np.zeros((B, 2, 4, 8), np.float32)), None, None, None, Tensor(inp)).numpy()
    seen = {}

    def sun_stub(x, plz, training):
        seen["x"], seen["plz"] = x.numpy(), plz.numpy()
        return Tensor(x.numpy() * np.float32(2.0)), None, None
    m.sun = sun_stub
    ldr = rng.uniform(0, 1, (B, H, W, 3)).astype(np.float32)
    cam1 = rng.uniform(0, 1, (B, H, W, 1)).astype(np.float32)
    cam2 = rng.uniform(0, 1, (B, H // 2, W // 2, 1)).astype(np.float32)
    cam3 = rng.uniform(0, 1, (B, H // 4, W // 4, 1)).astype(np.float32)
    pred = rng.uniform(0, 0.01, (B, H, W, 1)).astype(np.float32)
    rad_t, _, _ = G.model.sun_rad_estimation(m, Tensor(ldr), Tensor(cam1), Tensor(cam2), Tensor(cam3), Tensor(pred), False)
    out.update(sre_ldr=ldr, sre_cam1=cam1, sre_cam2=cam2, sre_cam3=cam3, sre_pred=pred, sre_x=seen["x"], sre_plz=seen["plz"],
               sre_out=rad_t.numpy())
    out["gen_blend"] = G.model.blending(m, Tensor(conv_out), Tensor(inp)).numpy()
    return out


def gradcam_vectors():
    """grad_cam.layer (grad_cam.py:29-45) from the reference's own file, with tf.gradients replaced by a recorded gradient."""
    tf = sys.modules["tensorflow"]
    rng = np.random.default_rng(13)
    A = np.maximum(rng.standard_normal((2, 4, 6, 5)), 0).astype(np.float32)
    grad = rng.standard_normal(A.shape).astype(np.float32)
    tf.gradients = lambda y, x: [Tensor(grad)]
    tf.reduce_mean = lambda x, axis=None, **k: Tensor(_to_np(x).mean(axis=axis, dtype=np.float32).astype(np.float32))
    tf.einsum = lambda eq, a, b: Tensor(np.einsum(eq.replace(" ", ""), _to_np(a), _to_np(b)).astype(np.float32))
    C = load("/root/reference/grad_cam.py", "ref_grad_cam")
    cam = C.layer(Tensor(np.zeros(2, np.float32)), Tensor(A)).numpy()
    return {"cam_A": A, "cam_grad": grad, "cam_out": cam}


def sunpose_order():
    """Call order of sunpose_net.model.sunposeEstimation (sunpose_net.py:54-72) from the reference's own file: every layer is a
    stand-in that appends its name to a log and tags the tensor, so the golden records the wiring (which tensors are returned as
    activation maps, where the pools sit, flatten -> fc1 -> relu -> fc2 -> relu -> softmax)."""
    tf = sys.modules["tensorflow"]
    log = []

    def stage(name):
        def f(x, *a, **k):
            log.append(name)
            return x
        return f
    tf.nn.softmax = lambda x: (log.append("softmax"), x)[1]
    ops = types.ModuleType("ops")
    ops.conv2d = ops.relu = ops.maxpool2d = lambda *a, **k: None
    sys.modules["ops"] = ops
    sys.modules["distortion_aware_ops"] = types.ModuleType("distortion_aware_ops")
    P = load("/root/reference/sunpose_net.py", "ref_sunpose_net")
    m = object.__new__(P.model)
    for name in ("sunlayer1", "pool1_s", "sunlayer2", "pool2_s", "sunlayer3", "pool3_s", "flat", "fc1", "actv1_s", "fc2", "actv2_s"):
        setattr(m, name, stage(name))
    x = Tensor(np.zeros((1, 2, 2, 1), np.float32))
    sm, acts = P.model.sunposeEstimation(m, x, False)
    lay = object.__new__(P.sunposeLayer)
    for name in ("conv1", "norm1", "actv1", "conv2", "norm2", "actv2"):
        setattr(lay, name, stage("layer." + name))
    P.sunposeLayer.call(lay, x, False)
    return {"sunpose_call_order": np.array(log), "sunpose_n_acts": np.array([len(acts)])}


def main():
    install()
    U = load("/root/reference/tf_utils.py", "ref_tf_utils")
    S = load("/root/reference/sunrad_net.py", "ref_sunrad_net")
    rng = np.random.default_rng(0)
    out = {}
    # log codec (tf_utils.py:263-280)
    x = (rng.uniform(0, 1, (2, 4, 8, 3)) ** 4 * 100).astype(np.float32)
    out["codec_x"] = x
    out["codec_compressed"] = U.hdr_logCompression(Tensor(x)).numpy()
    out["codec_roundtrip"] = U.hdr_logDecompression(Tensor(out["codec_compressed"])).numpy()
    # camera response lookup (tf_utils.py:191-255)
    K = 1024
    crf = (np.linspace(0, 1, K)[None, :] ** (1 / rng.uniform(1.5, 3, (3, 1)))).astype(np.float32)
    xr = rng.uniform(0, 1, (3, 5, 7, 3)).astype(np.float32)
    xr[0, 0, 0] = [0.0, 1.0, 0.5]
    out["rf_crf"], out["rf_x"] = crf, xr
    out["rf_y"] = U.apply_rf(Tensor(xr), Tensor(crf)).numpy()
    # sun-position bins and directions (tf_utils.py:95-129)
    h, w = 8, 32
    out["bins_hw"] = np.array([h, w])
    out["bins"] = np.stack([U.sunpose_init(Tensor(np.float32(i)), h, w).numpy() for i in range(h * w)])
    pts = np.array([[15.0, 2.5], [0.0, 0.0], [31.0, 7.9], [7.25, 4.0]], np.float32)
    out["s2w_pts"] = pts
    out["s2w"] = np.stack([U.sphere2world((Tensor(p[0]), Tensor(p[1])), h, w, skydome=True).numpy() for p in pts])
    # arithmetic tail of sunRadNet.call (sunrad_net.py:56-71) with the conv stack replaced by fixed head outputs
    net = S.sunRadNet()
    B, H, W = 3, 4, 8
    heads = rng.standard_normal((B, 2)).astype(np.float32) * 2
    ident = lambda t, *a: t
    net.d1 = net.d2 = net.d3 = net.d4 = ident
    net.flat = ident
    net.gamma = lambda t: Tensor(heads[:, 0:1])
    net.beta = lambda t: Tensor(heads[:, 1:2])
    sm = rng.uniform(0, 1, (B, H, W, 1)).astype(np.float32)
    sm[0, 0, 0, 0] = 1.0                                     # the normalised maximum: the peak of the radiance function
    rad, g_in, b_in = net.call(Tensor(sm), Tensor(np.zeros((B, 1), np.float32)), False)
    out["rad_heads"], out["rad_x"], out["rad_y"] = heads, sm, rad.numpy()
    out["rad_gamma_in"], out["rad_beta_in"] = g_in.numpy(), b_in.numpy()
    out.update(train_vectors())
    out.update(generator_vectors())
    out.update(gradcam_vectors())
    out.update(sunpose_order())
    np.savez_compressed(os.path.join(HERE, "utils_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
