"""Per-stage error report of generator inference against the oracle (run on a GPU box): python tools/dbg_inference.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
from oracle import model_oracle as M

pkg = load_package()
rel = lambda a, b: float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(np.asarray(b, np.float64)), 1e-30))
for da in (True, False):
    rng = np.random.default_rng(4)
    B, H, W = 2, 32, 128
    ldr = (np.round(255 * rng.uniform(0, 1, (B, H, W, 3))) / 255).astype(np.float32)
    wg, ws = M.random_full_generator_weights(3, H, W), M.random_sunpose_weights(5, H, W)
    gen, sun = pkg.inference.build_models(batch_size=B, im_height=H, im_width=W, distortion_aware_sunpose=da)
    x = torch.from_numpy(ldr).cuda()
    sun.sunposeEstimation(x); gen.set_weights(wg); sun.set_weights(ws)
    d = M.generator_inference(ldr, wg, ws, distortion_aware_sunpose=da, acc_dtype=torch.float64, details=True)
    res = gen.encode(x)
    sky = gen.sky_decode(res, x)
    sm, acts = sun.sunposeEstimation(x)
    yc = sun.class_score(sm)
    cams = [pkg.grad_cam.layer(yc, a) for a in acts]
    rad, _, _ = gen.sun_rad_estimation(x, *cams, sm, training=False, log_compress=True)
    sun_gamma = gen.sun_decode(res, *cams, rad)
    fin = gen.sun_decode(res, *cams, rad, blend_with=sky, log_decompress=True)
    torch.cuda.synchronize()
    print(f"--- distortion_aware_sunpose={da}")
    print("sky_gamma", rel(sky.cpu(), d["sky_gamma"]), "sm", rel(sm.cpu(), d["sm"]))
    for i in range(3):
        print(f"act{i+1}", rel(acts[i].cpu(), d["acts"][i]), f"cam{i+1}", rel(cams[i].cpu(), d["cams"][i]), "absmax", float(d["cams"][i].max()))
    print("sun_rad_gamma", rel(rad.cpu(), d["sun_rad_gamma"]), "sun_gamma", rel(sun_gamma.cpu(), d["sun_gamma"]))
    logl = lambda y: np.log1p(10 * np.asarray(y, np.float64)) / np.log(11.0)
    print("final lin", rel(fin.cpu(), d["y_lin"]), "final logl", rel(logl(fin.cpu().numpy()), logl(d["y_lin"].numpy())))
    full = pkg.inference.generator_in_step(gen, sun, x)
    print("one-call vs staged max abs diff", float((full - fin).abs().max()))
