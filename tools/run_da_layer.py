"""Runs one distortion-aware conv layer a few times (for ncu): python tools/run_da_layer.py B h w C F k [mode] [reps] [path]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("hdr-map-reconstruction-from-a-single-ldr-sky-panoramic-image-for-outdoor-illumination-estimation_b200")
B, h, w, C, F, k = (int(v) for v in sys.argv[1:7])
mode = sys.argv[7] if len(sys.argv) > 7 else "tf32"
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 5
path = sys.argv[9] if len(sys.argv) > 9 else None
torch.manual_seed(0)
x = torch.randn(B, h, w, C, device="cuda")
layer = pkg.conv2d(F, kernel_size=k, math_mode=mode)
layer.build((B, h, w, C))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
for s, e in ev:
    flush.zero_()
    s.record()
    y = layer(x, kernel_path=path)
    e.record()
torch.cuda.synchronize()
ms = [s.elapsed_time(e) for s, e in ev]
fl = 2.0 * B * h * w * k * k * C * F
print("layer", (B, h, w, C, F, k), mode, path, "ms", [round(m, 4) for m in ms], "TFLOP/s best %.1f" % (fl / min(ms) / 1e9))
