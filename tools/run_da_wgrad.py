"""Runs the distortion-aware weight gradient a few times (for ncu): python tools/run_da_wgrad.py B h w C F k [reps]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("hdr-map-reconstruction-from-a-single-ldr-sky-panoramic-image-for-outdoor-illumination-estimation_b200")
D = pkg.distortion_aware_ops
B, h, w, C, F, k = (int(v) for v in sys.argv[1:7])
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 5
torch.manual_seed(0)
x = torch.randn(B, h, w, C, device="cuda")
dy = torch.randn(B, h, w, F, device="cuda")
layer = pkg.conv2d(F, kernel_size=k, math_mode="tf32")
layer.build((B, h, w, C))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
fl = 2.0 * B * h * w * k * k * C * F
ms = []
for _ in range(reps):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    D.conv2d_backward(layer, x, dy, need_dx=False)
    e.record()
    torch.cuda.synchronize()
    ms.append(s.elapsed_time(e))
print("wgrad", (B, h, w, C, F, k), "ms", [round(m, 4) for m in ms], "TFLOP/s best %.1f" % (fl / min(ms[1:] or ms) / 1e9))
# the pipelined kernel of csrc/conv_bwd.cu through sky_conv2d_bwd_filter (offsets != NULL)
lib, check = pkg._lib.LIB, pkg._lib.check
dk, db = torch.empty_like(layer.kernel), torch.empty_like(layer.bias)
st = lambda: torch.cuda.current_stream().cuda_stream
ms = []
for _ in range(reps):
    flush.zero_()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    check(lib.sky_conv2d_bwd_filter(x.data_ptr(), dy.data_ptr(), layer.offset_table.data_ptr(), dk.data_ptr(), db.data_ptr(), B, h, w, C, C, F, k, 1, 0, st()))
    e.record()
    torch.cuda.synchronize()
    ms.append(s.elapsed_time(e))
print("wgrad (pipelined kernel, sky_conv2d_bwd_filter)", (B, h, w, C, F, k), "ms", [round(m, 4) for m in ms], "TFLOP/s best %.1f" % (fl / min(ms[1:] or ms) / 1e9))
