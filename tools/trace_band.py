"""Timeline of one CTA of the band-staged conv kernel at the bench shape (development aid).
SKY_DEBUG_FLAGS=2097152 python tools/trace_band.py"""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
B, h, w, C, F, k = 32, 8, 32, 128, 128, 3
layer = pkg.conv2d(F, kernel_size=k)
x = torch.randn(B, h, w, C, device="cuda")
layer.build(tuple(x.shape))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
names = ["start", "prologue", "table0", "band0", "prod_done", "acc_ready", "epi_done", "end"]
for rep in range(4):
    flush.zero_()
    layer(x)
    torch.cuda.synchronize()
    ts = np.zeros(16, np.uint64)
    pkg._lib.check(pkg._lib.LIB.sky_debug_band_trace(ts.ctypes.data))
    t0 = int(ts[0])
    print(" ".join(f"{n}={(int(ts[i]) - t0) / 1000:.1f}us" for i, n in enumerate(names)))
    print("   group0 atoms 16,20,24,28 (free,filled):", " ".join(f"{(int(ts[i]) - t0) / 1000:.2f}" for i in range(8, 16)))
