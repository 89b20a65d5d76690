"""Debug: strip kernel vs the direct / scatter / band kernels at B = 32 on the train step's layer shapes."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("hdr-map-reconstruction-from-a-single-ldr-sky-panoramic-image-for-outdoor-illumination-estimation_b200")
D, L = pkg.distortion_aware_ops, pkg._lib
lib, check = L.LIB, L.check
st = lambda: torch.cuda.current_stream().cuda_stream
def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
torch.manual_seed(0)
B = 32
# DA forward / data gradient
for (h, w, C, F, k) in ((32, 128, 32, 32, 7), (16, 64, 64, 64, 3), (8, 32, 128, 128, 3), (16, 64, 32, 64, 3)):
    x = torch.randn(B, h, w, C, device="cuda"); dy = torch.randn(B, h, w, F, device="cuda")
    layer = pkg.conv2d(F, kernel_size=k, math_mode="tf32"); layer.build((B, h, w, C)); layer.bias.normal_()
    s1 = torch.zeros(B, F, 2, dtype=torch.float64, device="cuda"); s2 = torch.zeros_like(s1)
    a = layer(x, stats=s1); b = layer(x, stats=s2, kernel_path="band")
    D.DA_BACKWARD_KERNEL = "strip"; g1 = D.conv2d_backward(layer, x, dy, need_dw=False)[0]
    D.DA_BACKWARD_KERNEL = "scatter"; g2 = D.conv2d_backward(layer, x, dy, need_dw=False)[0]
    D.DA_BACKWARD_KERNEL = "strip"
    base = torch.randn_like(x); acc = base.clone()
    D.conv2d_backward(layer, x, dy, need_dw=False, dx_out=acc, accumulate_dx=True)
    print("DA", (h, w, C, F, k), "fwd strip/band %.2e stats %.2e | dgrad strip/scatter %.2e accumulate %.2e" % (rel(a, b), rel(s1, s2), rel(g1, g2), rel(acc - base, g2)))
# plain convs through ops.conv2d
for (h, w, C, F, k, s) in ((32, 128, 32, 3, 7, 1), (32, 128, 64, 32, 3, 1), (16, 64, 64, 128, 4, 2), (8, 32, 128, 256, 4, 2), (32, 128, 32, 64, 3, 2), (4, 16, 256, 512, 4, 1), (16, 16, 256, 256, 3, 1)):
    x = torch.randn(B, h, w, C, device="cuda")
    conv = pkg.ops.conv2d(output_channels=F, k_h=k, k_w=k, strides=s, math_mode="tf32")
    y = conv(x)
    packed = conv._packed_weights()
    y2 = torch.empty_like(y)
    check(lib.sky_conv2d_fwd(x.data_ptr(), packed.data_ptr(), conv._bias().data_ptr(), y2.data_ptr(), None, None, B, h, w, C, F, k, s, L.EPI_FORCE_DIRECT, 0.0, L.MATH_TF32, st()))
    dy = torch.randn_like(y)
    dx = conv.backward_data(x, dy)
    tp = conv._tpack
    dx2 = torch.empty_like(x)
    if F > 4:
        check(lib.sky_conv2d_bwd_data(dy.data_ptr(), tp.packed.data_ptr(), dx2.data_ptr(), None, B, h, w, C, F, k, s, L.EPI_FORCE_DIRECT, 0.0, L.MATH_TF32, st()))
    else:
        dx2 = dx
    print("plain", (h, w, C, F, k, s), "fwd strip/direct %.2e | dgrad strip/direct %.2e" % (rel(y, y2), rel(dx, dx2)))
