"""Debug: backward of sunlayer3 step by step, forward by the strip kernel vs the band kernel, same upstream gradient."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import model_oracle as M
pkg = importlib.import_module("hdr-map-reconstruction-from-a-single-ldr-sky-panoramic-image-for-outdoor-illumination-estimation_b200")
D = pkg.distortion_aware_ops
LIB, check, _stream = pkg._lib.LIB, pkg._lib.check, D._stream
rng = np.random.default_rng(2)
B, H, W = 2, 32, 128
ldr = (np.round(255 * rng.uniform(0, 1, (B, H, W, 3))) / 255).astype(np.float32)
ws = M.random_sunpose_weights(seed=5, H=H, W=W)
g_out = torch.from_numpy(rng.standard_normal((B, 8, 32, 128)).astype(np.float32)).cuda()
res = {}
for path in ("strip", "band"):
    D.DA_FORWARD_KERNEL = path
    net = pkg.sunpose_net.model(im_height=H, im_width=W, distortion_aware=True, math_mode="3xtf32")
    x = torch.from_numpy(ldr).cuda()
    net.sunposeEstimation(x)
    net.set_weights(ws)
    net.sunposeEstimation(x, training=True)
    layer = net.sunlayer3
    xin, conv1, actv1, conv2, actv2 = layer._saved
    Bq, h, w, F = conv2.shape
    sums = torch.empty(Bq, F, 2, dtype=torch.float64, device="cuda")
    dg, db = torch.zeros(F, device="cuda"), torch.zeros(F, device="cuda")
    def norm_bwd(norm, pre, stats, dy, act):
        dx = torch.empty_like(pre)
        dg.zero_(); db.zero_()
        check(LIB.sky_instnorm_bwd(pre.data_ptr(), stats.data_ptr(), norm.gamma.data_ptr(), dy.data_ptr(), act.data_ptr(), None,
                                   sums.data_ptr(), dx.data_ptr(), dg.data_ptr(), db.data_ptr(), Bq, h, w, F, norm.epsilon, 0.0, _stream()))
        return dx
    g1 = norm_bwd(layer.norm2, conv2, layer._stats[1], g_out, actv2)
    g2 = D.conv2d_backward(layer.conv2, actv1, g1, need_dw=False)[0]
    g3 = norm_bwd(layer.norm1, conv1, layer._stats[0], g2, actv1)
    res[path] = dict(conv1=conv1.clone(), actv1=actv1.clone(), conv2=conv2.clone(), actv2=actv2.clone(), s0=layer._stats[0].clone(),
                     s1=layer._stats[1].clone(), g1=g1.clone(), g2=g2.clone(), g3=g3.clone(), db=db.clone(), dg=dg.clone())
def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))
for k in res["strip"]:
    a, b = res["strip"][k], res["band"][k]
    print("%-6s rel %.3e  mean diff %.3e  mean |b| %.3e" % (k, rel(a, b), float((a.double() - b.double()).mean()), float(b.double().abs().mean())))
a, b = res["strip"]["actv1"], res["band"]["actv1"]
print("mask flips", int(((a > 0) != (b > 0)).sum()), "of", a.numel())
a, b = res["strip"]["actv2"], res["band"]["actv2"]
print("mask2 flips", int(((a > 0) != (b > 0)).sum()), "of", a.numel())
