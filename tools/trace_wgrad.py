"""Timeline of CTA 0 of the strip weight-gradient kernel: python tools/trace_wgrad.py B h w C F k"""
import ctypes, importlib, os, sys
os.environ["SKY_WGRAD_TRACE"] = "1"
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("hdr-map-reconstruction-from-a-single-ldr-sky-panoramic-image-for-outdoor-illumination-estimation_b200")
D = pkg.distortion_aware_ops
B, h, w, C, F, k = (int(v) for v in sys.argv[1:7])
x, dy = torch.randn(B, h, w, C, device="cuda"), torch.randn(B, h, w, F, device="cuda")
layer = pkg.conv2d(F, kernel_size=k, math_mode="tf32")
layer.build((B, h, w, C))
for _ in range(3):
    D.conv2d_backward(layer, x, dy, need_dx=False)
torch.cuda.synchronize()
buf = np.zeros(128, np.uint64)
pkg._lib.check(pkg._lib.LIB.sky_debug_wgrad_trace(buf.ctypes.data))
t0 = int(buf[102])
rel = lambda v: (int(v) - t0) / 1e3 if v else float("nan")
print("kernel", rel(buf[103]), "us; first drain", rel(buf[100]), "->", rel(buf[101]))
for i in range(12):
    print("item %2d  mma: wait %8.2f got %8.2f issued %8.2f | producer: stage free %8.2f strip done %8.2f dy done %8.2f" %
          (i, rel(buf[4 * i]), rel(buf[4 * i + 1]), rel(buf[4 * i + 2]), rel(buf[48 + 4 * i]), rel(buf[49 + 4 * i]), rel(buf[50 + 4 * i])))
print("every 32nd item issued at (us):", [round(rel(buf[104 + i]), 1) for i in range(24) if buf[104 + i]])
