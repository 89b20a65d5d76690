"""Kernel-level sweep of the distortion-aware conv (BASELINE config 4): CUDA-event timing, L2 flushed between reps.
Usage: python tools/bench_conv.py [--sizes 32x128,64x256,128x512] [--batch 64] [--reps 10] [--only fwd] [--quick]"""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

LAYERS = [(128, 128, 3), (64, 64, 3), (32, 32, 7), (3, 32, 7), (32, 3, 7)]


def time_op(fn, reps, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="32x128,64x256,128x512")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--math", default="tf32")
    ap.add_argument("--layers", default="all")
    ap.add_argument("--bwd", action="store_true", help="also time dgrad and wgrad")
    ap.add_argument("--trunk-scale", action="store_true", help="run C128 layers at H/4 x W/4 (their site in the model)")
    args = ap.parse_args()
    pkg = load_package()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    layers = LAYERS if args.layers == "all" else [tuple(int(v) for v in l.split(":")) for l in args.layers.split(",")]
    for size in args.sizes.split(","):
        H, W = (int(v) for v in size.split("x"))
        for (C, F, k) in layers:
            h, w = (H // 4, W // 4) if (args.trunk_scale and C == 128) else (H, W)
            B = args.batch
            x = torch.randn(B, h, w, C, device="cuda")
            layer = pkg.conv2d(F, kernel_size=k, math_mode=args.math)
            layer.build(tuple(x.shape))
            ms = time_op(lambda: layer(x), args.reps, flush)
            M = B * h * w
            flops = 2.0 * M * k * k * C * F
            bytes_ = 4.0 * (M * C + M * F + k * k * C * F + F) + 8 * h * k * k
            print(json.dumps(dict(op="da_conv2d_fwd", math=args.math, B=B, h=h, w=w, C=C, F=F, k=k, ms=round(ms, 4),
                                  tflops=round(flops / ms / 1e9, 2), gbs=round(bytes_ / ms / 1e6, 1),
                                  frac_tf32_peak=round(flops / ms / 1e9 / (peaks["bf16_tflops"] / 2), 4),
                                  frac_hbm=round(bytes_ / ms / 1e6 / peaks["hbm_gbs"], 4))), flush=True)
            if args.bwd and C % 32 == 0 and F % 32 == 0:
                dy = torch.randn(B, h, w, F, device="cuda")
                bw = pkg.distortion_aware_ops.conv2d_backward
                for name, fn in (("da_conv2d_bwd_data", lambda: bw(layer, x, dy, need_dw=False)),
                                 ("da_conv2d_bwd_filter", lambda: bw(layer, x, dy, need_dx=False))):
                    ms = time_op(fn, args.reps, flush)
                    print(json.dumps(dict(op=name, math="tf32", B=B, h=h, w=w, C=C, F=F, k=k, ms=round(ms, 4),
                                          tflops=round(flops / ms / 1e9, 2),
                                          frac_tf32_peak=round(flops / ms / 1e9 / (peaks["bf16_tflops"] / 2), 4))), flush=True)
                del dy
            del x, layer


if __name__ == "__main__":
    main()
