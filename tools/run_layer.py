"""Run one layer shape a few times (for ncu captures):  python tools/run_layer.py dense B K N | da B h w C F k | plain B h w C F k stride"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
kind, a = sys.argv[1], [int(v) for v in sys.argv[2:]]
torch.manual_seed(0)
if kind == "dense":
    B, K, N = a
    d = pkg.sunpose_net.Dense(N); d.build((B, K))
    x = torch.randn(B, K, device="cuda")
    fn = lambda: d(x, relu=True)
elif kind == "da":
    B, h, w, C, F, k = a
    layer = pkg.conv2d(F, kernel_size=k)
    x = torch.randn(B, h, w, C, device="cuda")
    fn = lambda: layer(x)
elif kind == "dgrad":
    B, h, w, C, F, k = a
    layer = pkg.conv2d(F, kernel_size=k)
    x = torch.randn(B, h, w, C, device="cuda")
    layer(x)
    dy = torch.randn(B, h, w, F, device="cuda")
    fn = lambda: pkg.distortion_aware_ops.conv2d_backward(layer, x, dy, need_dw=False)
elif kind == "wgrad":
    B, h, w, C, F, k = a
    layer = pkg.conv2d(F, kernel_size=k)
    x = torch.randn(B, h, w, C, device="cuda")
    layer(x)
    dy = torch.randn(B, h, w, F, device="cuda")
    fn = lambda: pkg.distortion_aware_ops.conv2d_backward(layer, x, dy, need_dx=False)
else:
    B, h, w, C, F, k, s = a
    layer = pkg.ops.conv2d(output_channels=F, k_h=k, k_w=k, strides=s)
    x = torch.randn(B, h, w, C, device="cuda")
    fn = lambda: layer(x)
for _ in range(3):
    fn()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(10):
    fn()
ev[1].record(); torch.cuda.synchronize()
print(kind, a, "avg ms", ev[0].elapsed_time(ev[1]) / 10)
