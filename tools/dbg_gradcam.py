"""Raw Grad-CAM gradients vs oracle autograd (run on a GPU box)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
from oracle import model_oracle as M
pkg = load_package()
rel = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(np.asarray(b, np.float64)), 1e-30))
rng = np.random.default_rng(1)
B, H, W = 2, 32, 128
ldr = (np.round(255 * rng.uniform(0, 1, (B, H, W, 3))) / 255).astype(np.float32)
w = M.random_sunpose_weights(seed=5, H=H, W=W)
xt = torch.from_numpy(ldr).double().requires_grad_(True)
want_sm, want_acts = M.sunpose_estimation(xt, w, distortion_aware=True, acc_dtype=torch.float64)
want_g = torch.autograd.grad(want_sm.amax(dim=1).sum(), want_acts)
for mode in ("tf32", "3xtf32"):
    net = pkg.sunpose_net.model(im_height=H, im_width=W, distortion_aware=True, math_mode=mode)
    x = torch.from_numpy(ldr).cuda()
    net.sunposeEstimation(x); net.set_weights(w)
    sm, acts = net.sunposeEstimation(x)
    yc = net.class_score(sm)
    print("mode", mode)
    for i, a in enumerate(acts):
        g = yc.gradient(a).cpu().numpy()
        wg = want_g[i].numpy()
        wmean_got, wmean_want = g.mean(axis=(1, 2)), wg.mean(axis=(1, 2))
        print(f" grad{i+1} rel {rel(g, wg):.3e}  |g| {np.abs(wg).mean():.3e}  channel-mean rel {rel(wmean_got, wmean_want):.3e}  "
              f"mean/abs-mean {np.abs(wmean_want).mean() / np.abs(wg).mean():.3e}  act rel {rel(a.cpu(), want_acts[i].detach()):.3e}")
        # exact-gradient cams from OUR activations isolate the activation error from the gradient error
        cam_mix = torch.relu(torch.einsum('bc,bwhc->bwh', torch.from_numpy(wmean_want), a.cpu().double()))
        cam_want = torch.relu(torch.einsum('bc,bwhc->bwh', torch.from_numpy(wmean_want), want_acts[i].detach()))
        cam_got = pkg.grad_cam.layer(yc, a).cpu().numpy()[..., 0]
        print(f"   cam rel {rel(cam_got, cam_want):.3e}   cam with exact weights but our activations {rel(cam_mix, cam_want):.3e}")
