"""Debug: error of the strip / band / direct forward kernels against the fp64 oracle (reference factors and row factors)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import da_oracle as O
pkg = importlib.import_module("hdr-map-reconstruction-from-a-single-ldr-sky-panoramic-image-for-outdoor-illumination-estimation_b200")
def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))
for (B, h, w, C, F, k) in [(2, 8, 32, 32, 16, 3), (2, 8, 32, 128, 64, 3), (2, 8, 32, 128, 128, 3), (2, 8, 32, 32, 128, 3), (8, 8, 32, 32, 16, 3), (2, 32, 128, 32, 16, 3)]:
    rng = np.random.default_rng(1)
    x = rng.standard_normal((B, h, w, C)).astype(np.float32)
    kern = (rng.standard_normal((k * k * C, F)) / np.sqrt(k * k * C)).astype(np.float32)
    b = rng.standard_normal(F).astype(np.float32)
    want = O.conv2d_forward(x, kern, b, k, acc_dtype=torch.float64).numpy()
    with O.row_factors():
        want_row = O.conv2d_forward(x, kern, b, k, acc_dtype=torch.float64).numpy()
    xd = torch.from_numpy(x).cuda()
    for mode in ("3xtf32",):
        layer = pkg.conv2d(F, kernel_size=k, kernel_initializer=kern, bias_initializer=b, math_mode=mode)
        s = layer(xd).cpu().numpy()
        bd = layer(xd, kernel_path="band").cpu().numpy()
        dr = layer(xd, force_direct=True).cpu().numpy()
        print((B, h, w, C, F, k), mode, "strip/ref %.2e strip/row %.2e band/ref %.2e direct/ref %.2e strip/band %.2e" %
              (rel(s, want), rel(s, want_row), rel(bd, want), rel(dr, want), rel(s, bd.astype(np.float64))))
