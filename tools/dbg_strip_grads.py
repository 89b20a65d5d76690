"""Debug: sun pre-train step gradients with the row-strip kernel vs the band-staged kernel, each run twice (determinism)."""
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
gt = torch.softmax(torch.from_numpy(rng.standard_normal((B, H * W)).astype(np.float32) * 4), -1).numpy()
ws = M.random_sunpose_weights(seed=5, H=H, W=W)
runs = {}
for path in ("strip", "strip", "band", "band"):
    D.DA_FORWARD_KERNEL = path
    net = pkg.sunpose_net.model(im_height=H, im_width=W, distortion_aware=True, math_mode="3xtf32")
    tr = pkg.train_sun.SunTrainer(net, B, H, W, lr=1e-4)
    net.set_weights(ws)
    tr.sun_train_step([None, torch.from_numpy(ldr).cuda()], torch.from_numpy(gt).cuda())
    torch.cuda.synchronize()
    g = {}
    for name in ("sunlayer3", "sunlayer2", "sunlayer1"):
        layer = getattr(net, name)
        for i, (conv, norm) in enumerate(((layer.conv1, layer.norm1), (layer.conv2, layer.norm2)), start=1):
            g[f"{name}.conv{i}.kernel"] = tr._g(conv, "kernel").clone()
            g[f"{name}.norm{i}.gamma"] = tr._g(norm, "gamma").clone()
            g[f"{name}.norm{i}.beta"] = tr._g(norm, "beta").clone()
    runs.setdefault(path, []).append(g)
def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))
for k in runs["strip"][0]:
    print("%-28s strip-vs-strip %.2e  band-vs-band %.2e  strip-vs-band %.2e" % (k, rel(runs["strip"][0][k], runs["strip"][1][k]),
          rel(runs["band"][0][k], runs["band"][1][k]), rel(runs["strip"][0][k], runs["band"][0][k])))
