"""Generate tests/golden/da_golden.npz by executing the reference's own distortion_aware_ops.py (read from
/root/reference, unmodified) over the numpy/glibc TensorFlow stand-in in tf_shim.py.

Run in the build container only:   python tests/golden/make_golden.py
(/root/reference does not exist on the GPU box; the committed .npz is what the tests read.)

What is pinned:
  off_*      offset tables  distortion(h, w)            [h, k*k, 2] fp32   (bit-exact target)
  idx_*      the four index tensors fed to tf.gather_nd  [h, w, k*k, 2] int32 (y, x) per corner, batch 0 (bit-exact)
  wgt_*      the bilinear weights w0..w3                 [4, h, w, k*k] fp32 (bit-exact target)
  x_/k_/b_/y_*  seeded input, kernel, bias and the layer output (tolerance target; matmul order unspecified)
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_shim  # noqa: E402

REF = "/root/reference/distortion_aware_ops.py"


def load_reference():
    tf_shim.install()
    spec = importlib.util.spec_from_file_location("ref_distortion_aware_ops", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_layer(ref, kind, B, h, w, C, F, k, dilation, skydome, seed, out_hw=None):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, h, w, C)).astype(np.float32)
    kern = (rng.standard_normal((k * k * C, F)) / np.sqrt(k * k * C)).astype(np.float32)
    bias = rng.standard_normal((F,)).astype(np.float32)
    tf_shim.WEIGHT_INIT["kernel"] = lambda shape: kern
    tf_shim.WEIGHT_INIT["bias"] = lambda shape: bias
    tf_shim.GATHER_LOG.clear()
    tf_shim.TRACE.clear()
    if kind == "conv":
        layer = ref.conv2d(F, kernel_size=k, strides=1, dilation_rate=dilation, skydome=skydome)
    else:
        layer = ref.deconv2d(F, kernel_size=k, strides=1, output_imshape=list(out_hw), dilation_rate=dilation,
                             skydome=skydome)
    y = layer(tf_shim.Tensor(x)).numpy()
    off = layer.offset.numpy()                      # [1, H, W, k2, 2]
    assert np.array_equal(off[0, :, :1] + 0 * off[0], off[0], equal_nan=True)         # replicated over w (:266-268)
    idx = np.stack([g[0, ..., 1:3] for g in tf_shim.GATHER_LOG])     # [4, H, W, k2, 2] (y, x), batch 0
    for g in tf_shim.GATHER_LOG:                     # tiled over batch (:79): every sample sees the same indices
        assert (g[..., 1:3] == g[:1, ..., 1:3]).all()
    wgt = np.stack([m[0] for m in tf_shim.TRACE["mul4"][:4]])         # [4, H, W, k2]
    out = dict(x=x, k=kern, b=bias, y=y, off=off[0, :, 0], idx=idx.astype(np.int32), wgt=wgt)
    if kind == "deconv":
        out["resized"] = tf_shim.TRACE["resized"].astype(np.float32)
    return out


def main():
    ref = load_reference()
    blob = {}
    cases = []
    # (name, kind, B, h, w, C, F, k, dilation, skydome, out_hw)
    layer_cases = [
        ("c_8x32_k3", "conv", 2, 8, 32, 8, 8, 3, 1, True, None),
        ("c_4x16_k3_ns", "conv", 1, 4, 16, 4, 6, 3, 1, False, None),
        ("c_16x64_k7", "conv", 1, 16, 64, 3, 5, 7, 1, True, None),
        ("c_8x32_k3_d2", "conv", 1, 8, 32, 4, 4, 3, 2, True, None),
        ("c_12x48_k5", "conv", 1, 12, 48, 2, 3, 5, 1, True, None),
        ("d_4x16_to_8x32_k3", "deconv", 2, 4, 16, 8, 4, 3, 1, True, (8, 32)),
        ("d_8x32_to_16x64_k3", "deconv", 1, 8, 32, 4, 4, 3, 1, True, (16, 64)),
    ]
    for i, (name, kind, B, h, w, C, F, k, dil, sky, ohw) in enumerate(layer_cases):
        r = run_layer(ref, kind, B, h, w, C, F, k, dil, sky, seed=100 + i, out_hw=ohw)
        for key, val in r.items():
            blob[f"{key}__{name}"] = val
        cases.append(f"{name}|{kind}|{B}|{h}|{w}|{C}|{F}|{k}|{dil}|{int(sky)}|{ohw[0] if ohw else 0}|{ohw[1] if ohw else 0}")
        print(name, "y", r["y"].shape, "idx range", r["idx"].min(), r["idx"].max())

    # offset tables alone for every panorama / trunk size the configs use (distortion() needs no input tensor)
    off_cases = []
    tf_shim.WEIGHT_INIT["kernel"] = lambda shape: np.zeros(shape, np.float32)
    tf_shim.WEIGHT_INIT["bias"] = lambda shape: np.zeros(shape, np.float32)
    for (h, w) in [(8, 32), (16, 64), (32, 128), (64, 256), (128, 512)]:
        for k in (3, 7):
            for sky in (True, False):
                if h * k * k > 128 * 9 and not sky:
                    continue
                layer = ref.conv2d(1, kernel_size=k, skydome=sky)
                off = layer.distortion(h, w, skydome=sky).numpy()[0, :, 0]
                name = f"{h}x{w}_k{k}_{'sky' if sky else 'full'}"
                blob[f"offonly__{name}"] = off
                off_cases.append(f"{name}|{h}|{w}|{k}|1|{int(sky)}")
                print("offsets", name, off.shape)
    # the reference's failure mode: a tap with x=z=0 raises "undefined coordinates" (:252) at tiny sizes
    try:
        ref.conv2d(1, kernel_size=3).distortion(2, 8)
        undefined = 0
    except Exception as e:                               # noqa: BLE001
        undefined = int("undefined coordinates" in str(e))
    blob["undefined_2x8_k3"] = np.array(undefined)
    # kernel_size=1 cannot be built by the reference: tf.squeeze (:234) collapses the single tap, and the per-tap
    # loop (:238-239) then indexes a scalar.  Our layer mirrors this by rejecting k=1.
    try:
        ref.conv2d(1, kernel_size=1).distortion(8, 32)
        k1_raises = 0
    except Exception:                                    # noqa: BLE001
        k1_raises = 1
    blob["k1_raises"] = np.array(k1_raises)
    # even kernel sizes: AssertionError at :188
    try:
        ref.conv2d(1, kernel_size=4).distortion(8, 32)
        even_raises = 0
    except AssertionError:
        even_raises = 1
    blob["even_k_asserts"] = np.array(even_raises)
    blob["cases"] = np.array(cases)
    blob["off_cases"] = np.array(off_cases)
    np.savez_compressed(os.path.join(HERE, "da_golden.npz"), **blob)
    print("wrote", os.path.join(HERE, "da_golden.npz"), os.path.getsize(os.path.join(HERE, "da_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
