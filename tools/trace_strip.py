"""Timeline of CTA (0,0) of the row-strip kernel (globaltimer stamps; SKY_STRIP_TRACE=1): python tools/trace_strip.py B h w C F k [mode]"""
import ctypes, importlib, os, sys
os.environ["SKY_STRIP_TRACE"] = "1"
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("hdr-map-reconstruction-from-a-single-ldr-sky-panoramic-image-for-outdoor-illumination-estimation_b200")
B, h, w, C, F, k = (int(v) for v in sys.argv[1:7])
mode = sys.argv[7] if len(sys.argv) > 7 else "tf32"
x = torch.randn(B, h, w, C, device="cuda")
layer = pkg.conv2d(F, kernel_size=k, math_mode=mode)
layer.build((B, h, w, C))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
names = {0: "kernel start", 1: "prologue done", 2: "MMA: row class staged", 3: "MMA: last MMA issued", 4: "epilogue: accumulator complete",
         5: "epilogue done", 6: "kernel end", 7: "loader: first weight tile requested"}
for rep in range(3):
    if rep == 2:
        flush.zero_()
    layer(x)
    torch.cuda.synchronize()
    st = np.zeros(64, np.uint64)
    pkg._lib.check(pkg._lib.LIB.sky_debug_strip_trace(st.ctypes.data_as(ctypes.c_void_p)))
    t0 = int(st[0])
    print("---- rep", rep, "(L2 flushed)" if rep == 2 else "(L2 warm)")
    ev = []
    for i in range(64):
        if st[i] == 0:
            continue
        if i in names: nm = names[i]
        elif 8 <= i < 40: nm = "MMA: window %d %s" % ((i - 8) // 2, "weights landed" if (i & 1) else "waits for weights")
        else: nm = "producer group %d strip #%d written" % ((i - 40) // 3, (i - 40) % 3)
        ev.append((int(st[i]) - t0, nm))
    for t, nm in sorted(ev):
        print("%8.2f us  %s" % (t / 1e3, nm))
