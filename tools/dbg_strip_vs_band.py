"""Debug: sun-position network forward with the row-strip kernel vs the band-staged kernel (saved activations, instance-norm moments)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import model_oracle as M
pkg = importlib.import_module("hdr-map-reconstruction-from-a-single-ldr-sky-panoramic-image-for-outdoor-illumination-estimation_b200")
D = pkg.distortion_aware_ops
rng = np.random.default_rng(2)
B, H, W = 2, 32, 128
ldr = (np.round(255 * rng.uniform(0, 1, (B, H, W, 3))) / 255).astype(np.float32)
ws = M.random_sunpose_weights(seed=5, H=H, W=W)
out = {}
for path in ("strip", "band"):
    D.DA_FORWARD_KERNEL = path
    net = pkg.sunpose_net.model(im_height=H, im_width=W, distortion_aware=True, math_mode="3xtf32")
    x = torch.from_numpy(ldr).cuda()
    net.sunposeEstimation(x)
    net.set_weights(ws)
    y = net.sunposeEstimation(x)
    rec = {}
    for ln in ("sunlayer1", "sunlayer2", "sunlayer3"):
        L = getattr(net, ln)
        if L._saved is not None:
            xin, c1, a1, c2, a2 = L._saved
            rec[ln] = dict(c1=c1.clone(), a1=a1.clone(), c2=c2.clone(), a2=a2.clone(), s0=L._stats[0].clone(), s1=L._stats[1].clone())
    out[path] = rec
def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())
for ln in out["strip"]:
    for k in out["strip"][ln]:
        a, b = out["strip"][ln][k], out["band"][ln][k]
        print(ln, k, "rel %.3e" % rel(a, b), "maxabs %.3e" % float((a.double() - b.double()).abs().max()))
