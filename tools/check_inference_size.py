"""Ad-hoc parity check of generator inference at another panorama size (B = 1): python tools/check_inference_size.py 64 256"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
from oracle import model_oracle as M

H, W = int(sys.argv[1]), int(sys.argv[2])
pkg = load_package()
rng = np.random.default_rng(4)
ldr = (np.round(255 * rng.uniform(0, 1, (1, H, W, 3))) / 255).astype(np.float32)
wg, ws = M.random_full_generator_weights(3, H, W), M.random_sunpose_weights(5, H, W)
gen, sun = pkg.inference.build_models(batch_size=1, im_height=H, im_width=W)
x = torch.from_numpy(ldr).cuda()
sun.sunposeEstimation(x); gen.set_weights(wg); sun.set_weights(ws)
got = pkg.inference.generator_in_step(gen, sun, x).cpu().numpy()
t = time.time()
want = M.generator_inference(ldr, wg, ws, acc_dtype=torch.float64).numpy()
logl = lambda y: np.log1p(10 * np.asarray(y, np.float64)) / np.log(11.0)
r = np.linalg.norm(logl(got) - logl(want)) / np.linalg.norm(logl(want))
print(f"generator inference 1x{H}x{W}: rel_l2 (log-luminance) = {r:.3e}, finite = {bool(np.isfinite(got).all())}, oracle {time.time() - t:.1f} s")
