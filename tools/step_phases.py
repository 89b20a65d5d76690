"""Phase boundaries of one eager train step on its three streams (CUDA events): which stream the critical path runs through.
python tools/step_phases.py [B H W mode]"""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("hdr-map-reconstruction-from-a-single-ldr-sky-panoramic-image-for-outdoor-illumination-estimation_b200")
B, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (32, 32, 128)
mode = sys.argv[4] if len(sys.argv) > 4 else "3xtf32"
rng = np.random.default_rng(0)
ldr = torch.from_numpy((np.round(255 * rng.uniform(0, 1, (B, H, W, 3))) / 255).astype(np.float32)).cuda()
hdr = torch.from_numpy((rng.uniform(0, 1, (B, H, W, 3)) ** 3 * 3).astype(np.float32)).cuda()
gt = torch.softmax(torch.from_numpy(rng.standard_normal((B, H * W)).astype(np.float32) * 4), -1).cuda()
step = pkg.train.Step(batch_size=B, im_height=H, im_width=W, math_mode=mode)
for _ in range(3):
    step.train_step([hdr, ldr], gt)
torch.cuda.synchronize()
for rep in range(2):
    step._marks = []
    step.train_step([hdr, ldr], gt)
    torch.cuda.synchronize()
    marks, step._marks = step._marks, None
t0 = marks[0][1]
for name, ev in sorted(marks, key=lambda m: t0.elapsed_time(m[1])):
    print("%8.3f ms  %s" % (t0.elapsed_time(ev), name))
