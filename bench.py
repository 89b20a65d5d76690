#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native distortion-aware convolution path.

Workload (config.workload, default `sun_train`): BASELINE.json configs[1], the sun pre-train step at batch 32, 32x128 =
train_sun.sun_train_step (train_sun.py:220-264): sun-position network forward (distortion-aware wiring of sunpose_net.py:11,16) ->
Grad-CAM x3 at the ground-truth class -> KLDivergence + DoG L1 -> backward through every layer -> (world > 1: gradient all-reduce,
the 201 MB Dense part overlapped with the conv backward) -> Adam.  At N = 1 the same line carries `inference`: BASELINE configs[0],
"generator inference, random-init weights, synthetic 32x128 LDR sky-dome panoramas, batch 32" = inference.generator_in_step
(inference.py:81-112), which `--workload inference` also times on its own (with its own e2e / roofline keys).
`--workload sky` times the sky branch alone (inference.py:84-86), `--workload trunk` the DA residual trunk alone,
`--workload trunk_train` the data-parallel train step of the trunk.

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, one process per GPU under torchrun)
    python bench.py --impl reference [--steps K] [--warmup W]      reference arm: the oracle's CPU restatement of the
                                                                   same path on the host cores (TensorFlow itself is
                                                                   not installable here)
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_BLOCKS, C, K_SIZE = 6, 128, 3
UNIT = "panoramas/s"
METRICS = {"trunk_train": "panoramas/sec (32x128, data-parallel train step of the DA residual trunk: fwd + L2 loss + bwd + 1 NCCL all-reduce + RMSprop)",
           "sun_train": "panoramas/sec (32x128, sun-position network pre-train step: train_sun.sun_train_step, fwd + Grad-CAM + KL/DoG loss + bwd + Adam)",
           "inference": "panoramas/sec (32x128, generator inference: inference.generator_in_step, sky + sun branch)",
           "sky": "panoramas/sec (32x128, generator inference, sky branch: encode -> DA res-trunk -> sky_decode -> log-decompress)",
           "trunk": "panoramas/sec (32x128, inference: DA residual trunk forward)"}


def make_weights(seed=0):
    """numpy-seeded weights with the reference's initialiser distributions (glorot_uniform, zero bias, gamma 1, beta 0)."""
    rng = np.random.default_rng(seed)
    lim = (6.0 / (K_SIZE * K_SIZE * C + C)) ** 0.5
    blocks = []
    for _ in range(N_BLOCKS):
        w = {}
        for i in (1, 2):
            w[f"conv{i}_kernel"] = rng.uniform(-lim, lim, (K_SIZE * K_SIZE * C, C)).astype(np.float32)
            w[f"conv{i}_bias"] = np.zeros(C, np.float32)
            w[f"norm{i}_gamma"] = np.ones(C, np.float32)
            w[f"norm{i}_beta"] = np.zeros(C, np.float32)
        blocks.append(w)
    return blocks


def make_input(batch, h, w, seed):
    # trunk input = output of leaky_relu(IN(conv3_d(.))) (generator.py:104-106): unit-variance, LeakyReLU-shaped
    x = np.random.default_rng(seed).standard_normal((batch, h, w, C)).astype(np.float32)
    return np.where(x > 0, x, 0.1 * x).astype(np.float32)


def workload_name(batch, H, W, workload="sky"):
    if workload == "sun_train":
        return (f"sun_train_step (train_sun.py:220-264): LDR [{batch},{H},{W},3] -> sunpose_net (distortion-aware convs, 2 Dense {H * W}) -> "
                f"Grad-CAM x3 at the ground-truth class -> KLDivergence + DoG L1 -> backward through every layer -> gradient all-reduce "
                f"(world > 1) -> Adam; random-init weights, B={batch}/GPU")
    if workload == "inference":
        return (f"generator_inference (inference.py:81-112): LDR [{batch},{H},{W},3] -> encode + 6 DA resBlocks -> sky_decode; "
                f"sunpose_net (DA convs, 2 Dense {H * W}) -> Grad-CAM x3 (backward sweep) -> sunRadNet -> sun_decode -> alpha blend -> "
                f"hdr_logDecompression; random-init weights, B={batch}/GPU")
    if workload == "trunk_train":
        return (f"res_trunk_train_step: 6 resBlocks with distortion-aware convs, forward + synthetic L2 objective + backward "
                f"(IN/LeakyReLU bwd, DA dgrad/wgrad/dbias) + one all-reduce of the flat gradient buffer (7.1 MB) + fused Keras RMSprop, "
                f"B={batch}/GPU on the {H // 4}x{W // 4}x{C} trunk map of {H}x{W} panoramas")
    if workload == "trunk":
        return (f"res_trunk_fwd: 6 resBlocks = 12 distortion-aware conv2d (128->128, k=3, TF32) + 12 instance norms, "
                f"B={batch}/GPU, {H}x{W} panoramas -> trunk map {H // 4}x{W // 4}x{C}")
    return (f"generator_sky_inference (inference.py:84-86): LDR [{batch},{H},{W},3] -> encode (7x7/32, 3x3s2/64, 3x3s2/128, IN, lrelu) "
            f"-> 6 resBlocks with distortion-aware convs (generator.py:14,18) -> sky_decode (2 resize-deconv, 7x7/3, +LDR, relu) "
            f"-> hdr_logDecompression; random-init weights, B={batch}/GPU")


def make_generator_weights(seed=0):
    """Sky-branch weights with the reference's initialiser distributions (Keras glorot_uniform, zero bias, gamma 1, beta 0)."""
    rng = np.random.default_rng(seed)
    w = {}

    def conv(name, kk, cin, cout):
        lim = (6.0 / (kk * kk * cin + kk * kk * cout)) ** 0.5
        w[name] = (rng.uniform(-lim, lim, (kk, kk, cin, cout)).astype(np.float32), np.zeros(cout, np.float32))

    conv("conv1_d", 7, 3, 32); conv("conv2_d", 3, 32, 64); conv("conv3_d", 3, 64, 128)
    conv("conv3_f", 3, 128, 64); conv("conv2_f", 3, 64, 32); conv("conv1_f", 7, 32, 3)
    for name, c in (("norm1_d", 32), ("norm2_d", 64), ("norm3_d", 128), ("norm3_f", 64), ("norm2_f", 32)):
        w[name] = (np.ones(c, np.float32), np.zeros(c, np.float32))
    w["res"] = make_weights(seed + 1)
    return w


def make_inference_weights(H, W, seed=0):
    """Full generator (sky + sun decoder + sunRadNet) and sun-position weights with the reference's initialiser distributions:
    glorot_uniform kernels and zero biases, gamma 1 / beta 0 norms, N(0, 0.02) sunRadNet convs with fresh BatchNormalization."""
    rng = np.random.default_rng(seed + 50)
    wg = make_generator_weights(seed)

    def conv(kk, cin, cout):
        lim = (6.0 / (kk * kk * cin + kk * kk * cout)) ** 0.5
        return (rng.uniform(-lim, lim, (kk, kk, cin, cout)).astype(np.float32), np.zeros(cout, np.float32))

    wg["conv3_u"], wg["conv2_u"], wg["conv1_u"] = conv(3, 128, 64), conv(3, 64, 32), conv(7, 32, 3)
    wg["norm3_u"], wg["norm2_u"] = (np.ones(64, np.float32), np.zeros(64, np.float32)), (np.ones(32, np.float32), np.zeros(32, np.float32))
    sun, cin = {}, 6
    for name, f, norm in (("d1", 64, False), ("d2", 128, True), ("d3", 256, True), ("d4", 512, True)):
        d = {"kernel": (0.02 * rng.standard_normal((4, 4, cin, f))).astype(np.float32)}
        if norm:
            d.update(gamma=np.ones(f, np.float32), beta=np.zeros(f, np.float32), moving_mean=np.zeros(f, np.float32),
                     moving_variance=np.ones(f, np.float32))
        sun[name], cin = d, f
    flat = (H // 8) * (W // 8) * 512
    lim = (6.0 / (flat + 1)) ** 0.5
    for head in ("gamma", "beta"):
        sun[head] = (rng.uniform(-lim, lim, (flat, 1)).astype(np.float32), np.zeros(1, np.float32))
    wg["sun"] = sun
    ws, cin = {}, 3
    for name, f, k in (("sunlayer1", 32, 7), ("sunlayer2", 64, 3), ("sunlayer3", 128, 3)):
        d, c = {}, cin
        for i in (1, 2):
            lim = (6.0 / (k * k * c + f)) ** 0.5
            d[f"conv{i}_kernel"] = rng.uniform(-lim, lim, (k * k * c, f)).astype(np.float32)
            d[f"conv{i}_bias"] = np.zeros(f, np.float32)
            d[f"norm{i}_gamma"], d[f"norm{i}_beta"] = np.ones(f, np.float32), np.zeros(f, np.float32)
            c = f
        ws[name], cin = d, f
    fc = H * W
    for name, kin in (("fc1", (H // 8) * (W // 8) * 128), ("fc2", fc)):
        lim = (6.0 / (kin + fc)) ** 0.5
        ws[name] = (rng.uniform(-lim, lim, (kin, fc)).astype(np.float32), np.zeros(fc, np.float32))
    return wg, ws


def make_sunpose_gt(batch, H, W, seed):
    """sunpose_gt = von Mises-Fisher bump (kappa = 80) over the H*W sky bins (train.py:42-52) at azimuth W/2 - 1
    (train.py:32) and a random elevation row in [H/8, H/2] — computed by the package's own mirror of train.vMF (dataset.py)."""
    from __graft_entry__ import load_package
    D = load_package().dataset
    rng = np.random.default_rng(seed)
    bins = D.sunpose_bins(H, W)
    return np.stack([D.vMF(W * 0.5 - 1, rng.uniform(H / 8, H / 2), H, W, bins=bins) for _ in range(batch)]).astype(np.float32)


def make_ldr(batch, H, W, seed):
    """LDR = round(255 u) / 255, u ~ U[0,1)  (mimics train.py:84-92)."""
    return (np.round(255 * np.random.default_rng(seed).uniform(0, 1, (batch, H, W, 3))) / 255).astype(np.float32)


def oracle_step_fn(args, sample):
    """CPU restatement of the selected workload on `sample` panoramas (returns a zero-argument callable)."""
    import torch
    from oracle import model_oracle as M
    if args.workload == "trunk_train":
        blocks = [{k: torch.from_numpy(v).requires_grad_(True) for k, v in b.items()} for b in make_weights()]
        x = torch.from_numpy(make_input(sample, args.height // 4, args.width // 4, seed=1))
        tgt = torch.from_numpy(make_input(sample, args.height // 4, args.width // 4, seed=2))
        params = [v for blk in blocks for v in blk.values()]
        ms = [torch.zeros_like(v) for v in params]

        def step():
            loss = ((M.res_layer(x, blocks, K_SIZE) - tgt) ** 2).mean()
            grads = torch.autograd.grad(loss, params)
            with torch.no_grad():
                for v, g, m in zip(params, grads, ms):
                    m.mul_(0.9).add_(0.1 * g * g)
                    v.sub_(1e-4 * g / (m.sqrt() + 1e-7))
        return step
    if args.workload == "trunk":
        blocks = [{k: torch.from_numpy(v) for k, v in b.items()} for b in make_weights()]
        x = torch.from_numpy(make_input(sample, args.height // 4, args.width // 4, seed=1))
        return lambda: M.res_layer(x, blocks, K_SIZE)
    ldr = make_ldr(sample, args.height, args.width, seed=1)
    if args.workload == "sun_train":
        _, ws = make_inference_weights(args.height, args.width)
        gt = make_sunpose_gt(sample, args.height, args.width, seed=2)
        state = {}

        def step():
            # autograd through the restated forward + loss, then Adam (train_sun.py:257-258) on every variable
            _, grads = M.sun_train_step_grads(ldr, gt, ws, acc_dtype=torch.float32, with_gradcam=True)
            t = state["t"] = state.get("t", 0) + 1
            lr_t = 1e-4 * (1 - 0.999 ** t) ** 0.5 / (1 - 0.9 ** t)
            for name in ws:
                items = list(ws[name].items()) if isinstance(ws[name], dict) else list(enumerate(ws[name]))
                for key, val in items:
                    g = grads[name][key].numpy()
                    m, v = state.setdefault((name, key), [np.zeros_like(g), np.zeros_like(g)])
                    m *= 0.9; m += 0.1 * g
                    v *= 0.999; v += 0.001 * g * g
                    val -= (lr_t * m / (np.sqrt(v) + 1e-7)).astype(np.float32)
        return step
    if args.workload == "inference":
        wg, ws = make_inference_weights(args.height, args.width)
        return lambda: M.generator_inference(ldr, wg, ws, K_SIZE)
    w = make_generator_weights()
    return lambda: M.sky_branch(ldr, w, K_SIZE)


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([v.strip() for v in out.split(",")])
            except Exception:       # noqa: BLE001  (nvidia-smi missing: report empty clocks rather than fail the bench)
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def run_reference(args):
    """Reference arm: the path's CPU restatement (oracle port; TF cannot be installed) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import model_oracle as M
    torch.set_num_threads(os.cpu_count())
    sample = min(args.batch, 8)                     # bounded sample of the B=32 batch per step
    step = oracle_step_fn(args, sample)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = sample / dt
    METRIC = METRICS[args.workload].replace("32x128", f"{args.height}x{args.width}")
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.batch, args.height, args.width, args.workload),
                       "note": "reference dataflow (materialised pad/gather/blend/matmul) restated on torch-CPU; not TensorFlow"},
            "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                             "sample": f"{sample} of {args.batch} panoramas per step, {args.steps} steps"},
            "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def measure_inference(pkg, args, rank, flush):
    """Device-resident throughput of full generator inference (BASELINE configs[0]) measured in the same process: CUDA-graph replay,
    L2 flushed between steps, CUDA events.  Reported next to the train-step line so one run carries both halves of the metric."""
    import torch
    B, H, W = args.batch, args.height, args.width
    gen, sun = pkg.inference.build_models(batch_size=B, im_height=H, im_width=W, math_mode=args.math)
    x = torch.from_numpy(make_ldr(B, H, W, seed=1 + rank)).cuda()
    sun.sunposeEstimation(x)
    wg, ws = make_inference_weights(H, W)
    gen.set_weights(wg)
    sun.set_weights(ws)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(2):
            pkg.inference.generator_in_step(gen, sun, x)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            pkg.inference.generator_in_step(gen, sun, x)
    torch.cuda.synchronize()
    for _ in range(3):
        graph.replay()
    evs = []
    for _ in range(args.steps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        graph.replay()
        e.record()
        evs.append((s, e))
    torch.cuda.synchronize()
    t = sum(s.elapsed_time(e) for s, e in evs) / args.steps
    return {"workload": workload_name(B, H, W, "inference"), "value": round(B / (t * 1e-3), 1), "unit": UNIT, "ms_per_step": round(t, 4),
            "steps": args.steps, "cuda_graph": True}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = load_package()

    B, H, W = args.batch, args.height, args.width
    h, w = H // 4, W // 4
    METRIC = METRICS[args.workload].replace("32x128", f"{args.height}x{args.width}")
    trainer = None
    extra_inputs = []
    if args.workload == "trunk_train":
        trunk = pkg.resLayer((C,) * N_BLOCKS, C, k_h=K_SIZE, k_w=K_SIZE, math_mode=args.math)
        trunk.build((B, h, w, C))
        trunk.set_weights(make_weights())
        trainer = pkg.trunk_train.TrunkTrainer(trunk, (B, h, w, C), lr=1e-4)
        target = torch.from_numpy(make_input(B, h, w, seed=100 + rank)).cuda()
        loss_buf = trainer._loss

        def forward(inp):
            trainer.train_step(inp, target)
            return loss_buf
        x_host = torch.from_numpy(make_input(B, h, w, seed=1 + rank)).pin_memory()
        # per res-block: fwd 2 conv + 2 IN; bwd 2 x (IN reduce + IN apply) + 2 dgrad + 2 wgrad + 2 dbias + 2 pack (re-pack after the
        # update); + loss + rmsprop
        launches_per_step = None
    elif args.workload == "trunk":
        trunk = pkg.resLayer((C,) * N_BLOCKS, C, k_h=K_SIZE, k_w=K_SIZE, math_mode=args.math)
        trunk.build((B, h, w, C))
        trunk.set_weights(make_weights())
        forward = trunk
        x_host = torch.from_numpy(make_input(B, h, w, seed=1 + rank)).pin_memory()   # each rank: its own shard of the batch
        launches_per_step = None
    elif args.workload == "sun_train":
        sun = pkg.sunpose_net.model(im_height=H, im_width=W, math_mode=args.math)
        trainer = pkg.train_sun.SunTrainer(sun, B, H, W, lr=1e-4)
        sun.set_weights(make_inference_weights(H, W)[1])
        gt_host = torch.from_numpy(make_sunpose_gt(B, H, W, seed=200 + rank)).pin_memory()
        gt_dev = gt_host.cuda()
        extra_inputs.append((gt_dev, gt_host))        # the step's second input: copied host -> device every e2e step
        x_host = torch.from_numpy(make_ldr(B, H, W, seed=1 + rank)).pin_memory()
        trunk = None

        def forward(inp):
            trainer.sun_train_step([None, inp], gt_dev)
            return trainer.loss
        launches_per_step = None
    elif args.workload == "inference":
        gen, sun = pkg.inference.build_models(batch_size=B, im_height=H, im_width=W, math_mode=args.math)
        x_host = torch.from_numpy(make_ldr(B, H, W, seed=1 + rank)).pin_memory()
        sun.sunposeEstimation(x_host.cuda())          # builds the lazily created layers
        wg, ws = make_inference_weights(H, W)
        gen.set_weights(wg)
        sun.set_weights(ws)
        trunk = gen.res
        forward = lambda inp: pkg.inference.generator_in_step(gen, sun, inp)
        launches_per_step = None                      # counted from the entry points one step calls (see below)
    else:
        gen = pkg.model(batch_size=B, im_height=H, im_width=W, da_kernel_size=K_SIZE, math_mode=args.math)
        gen.build(B)
        gen.set_weights(make_generator_weights())
        trunk = gen.res
        forward = gen.sky_inference
        x_host = torch.from_numpy(make_ldr(B, H, W, seed=1 + rank)).pin_memory()
        # encoder 3 conv + 3 IN; trunk 12 conv + 12 IN; decoder 2 resize + 3 conv + 2 IN
        launches_per_step = None
    x = x_host.cuda()
    y_host = torch.empty_like(forward(x).cpu()).pin_memory()
    h2d_bytes = x_host.numel() * x_host.element_size() + sum(h.numel() * h.element_size() for _, h in extra_inputs)
    # kernel launches of one step, from the C-ABI entry points it calls (weights are packed / transposed by now)
    pkg._lib.LIB.counts = {}
    n0 = pkg._lib.LIB.sky_launch_count()
    forward(x)
    counted = pkg._lib.LIB.sky_launch_count() - n0          # exact: the library counts every kernel it launches
    abi_calls = {k: v for k, v in pkg._lib.LIB.counts.items() if k != "sky_launch_count"}
    pkg._lib.LIB.counts = None
    if launches_per_step is None:
        launches_per_step = counted
    if world > 1 and trainer is not None:
        trainer.flat_w.copy_(trainer.flat_w)     # replicas start from identical weights (same numpy seed on every rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")             # > 126 MB L2

    # ---- the step: captured once into a CUDA graph (24 kernel launches + 12 memsets) ----
    use_graph = trainer is None                  # the train step (NCCL all-reduce inside) is launched eagerly
    side = torch.cuda.Stream()
    graph = None
    if use_graph:
        with torch.cuda.stream(side):
            for _ in range(2):
                y = forward(x)                   # eager warm-up (packs weights, sizes scratch)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                y = forward(x)
    else:
        for _ in range(2):
            y = forward(x)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        evs = []
        for _ in range(steps):
            flush.zero_()                        # L2 flush between iterations, outside the timed events
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        return [s.elapsed_time(e) for s, e in evs]

    def run_step():
        if graph is not None:
            graph.replay()
        else:
            forward(x)

    def step_device():
        run_step()

    def step_e2e():
        x.copy_(x_host, non_blocking=True)       # H2D of the step's input(s) from pinned memory
        for dev_t, host_t in extra_inputs:
            dev_t.copy_(host_t, non_blocking=True)
        run_step()
        y_host.copy_(y, non_blocking=True)       # D2H of the step's result (inference: HDR map; training: the loss)

    for _ in range(max(args.warmup, 3)):
        step_device()
    # keep warming for at least 0.4 s of wall time: clocks, the caching allocator's per-stream pools (the step uses side streams)
    # and NCCL's channels reach their steady state only after a few dozen steps; untimed, the same count on every rank
    torch.cuda.synchronize()
    t_w = time.perf_counter()
    step_device()
    torch.cuda.synchronize()
    extra = int(min(200, max(0, 0.4 / max(time.perf_counter() - t_w, 1e-4))))
    if world > 1:                                   # the train step contains a collective: every rank must run the same number of steps
        ex = torch.tensor([extra], dtype=torch.int64, device="cuda")
        dist.all_reduce(ex, op=dist.ReduceOp.MAX)
        extra = int(ex.item())
    for _ in range(extra):
        step_device()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms = timed(step_device, args.steps)
    barrier()
    for _ in range(3):
        step_e2e()
    barrier()
    ms_e2e = timed(step_e2e, args.steps)
    barrier()
    clocks = sampler.stop() if sampler else None

    # ---- dominant kernel: the band-staged DA conv, timed per launch with CUDA events on its stream ----
    conv_ms = []
    if args.workload == "sun_train":
        # sunlayer1.conv2 (32 -> 32, 7x7, full resolution): the largest single launch of the step, forward direction
        pc, pk, pf, pm = 32, 7, 32, B * H * W
        probe_layer = sun.sunlayer1.conv2
        xt = torch.randn(B, H, W, pc, device="cuda")
        stats = torch.zeros(B, pf, 2, dtype=torch.float64, device="cuda")
    else:
        pc, pk, pf, pm = C, K_SIZE, C, B * h * w
        probe_layer = trunk.sequence[0].conv1
        xt = torch.from_numpy(make_input(B, h, w, seed=7)).cuda()          # a trunk-shaped activation
        stats = torch.zeros(B, C, 2, dtype=torch.float64, device="cuda")
    for rep in range(12 + 3):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        probe_layer.call(xt, stats=stats)
        e.record()
        torch.cuda.synchronize()
        if rep >= 3:
            conv_ms.append(s.elapsed_time(e))
    conv_t = float(np.mean(conv_ms))

    tot = torch.tensor([sum(ms), sum(ms_e2e)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)           # max over ranks
    t_dev, t_e2e = (float(v) / args.steps for v in tot.tolist())
    in_sync = None
    if world > 1 and trainer is not None:
        # data-parallel sanity: after the timed steps every replica must hold the same weights (same init, averaged gradients)
        chk = trainer.flat_w.double().sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        in_sync = bool((hi - lo).abs().item() <= 1e-9 * max(1.0, abs(hi.item())))

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else None
        bf16_peak = peaks["bf16_tflops"] if peaks else 1590.0
        flops = 2.0 * pm * (pk * pk * pc) * pf
        achieved = flops / (conv_t * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("k7" if args.workload == "sun_train" else "trunk", {}).get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": round(world * B / (t_dev * 1e-3), 1), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(t_dev, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "tf32" if args.math == "tf32" else "f32(3xtf32)", "data": "synthetic",
            "config": {"workload": workload_name(B, H, W, args.workload), "global_batch": world * B, "parallelism": (f"batch shards x{world}, " + ("gradient all-reduce (NCCL) of the flat buffer, Dense part overlapped with the conv backward"
                                                                     if trainer is not None else "no collective")),
                       "l2": "256 MB buffer written between timed iterations (outside the events)", "cuda_graph": graph is not None},
            "e2e": {"value": round(world * B / (t_e2e * 1e-3), 1), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": y_host.numel() * y_host.element_size(), "ms_per_step": round(t_e2e, 4)},
            "gpu_launches": launches_per_step * args.steps, "abi_calls_per_step": abi_calls,
            "roofline": {"kernel": "da_conv2d_fwd_band_kernel (%d->%d, k=%d, M=%d)" % (pc, pf, pk, pm), "bound": "tensor",
                         "achieved": round(achieved, 2), "peak": round(bf16_peak / 2, 1), "unit": "TFLOP/s",
                         "frac": round(achieved / (bf16_peak / 2), 4), "traffic": traffic,
                         "peak_note": ("TF32 operands: peak = 1/2 x measured bf16 burst (%s)" % ("of measured" if peaks else "of fallback")),
                         "frac_of_bf16_peak": round(achieved / bf16_peak, 4), "ms_per_launch": round(conv_t, 5),
                         "flops_per_launch": flops},
            "clocks": clocks,
        }
        if in_sync is not None:
            line["replicas_in_sync"] = in_sync
        if world == 1:
            if args.workload == "sun_train" and not args.no_inference:
                line["inference"] = measure_inference(pkg, args, rank, flush)
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args):
    """The oracle (a port of the reference dataflow) timed on the host cores, on a bounded sample."""
    import torch
    torch.set_num_threads(os.cpu_count())
    B = args.batch
    sample = min(B, 8)
    step = oracle_step_fn(args, sample)
    step()
    reps, t0 = 0, time.perf_counter()
    while reps < 3 or (time.perf_counter() - t0 < 10 and reps < 40):
        step()
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    return {"value": round(sample / dt, 3), "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": f"{sample} of {B} panoramas per step, {reps} steps, torch-CPU fp32 restatement of the reference dataflow"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="panoramas per GPU per step")
    ap.add_argument("--height", type=int, default=32)
    ap.add_argument("--width", type=int, default=128)
    ap.add_argument("--math", default="tf32", choices=["tf32", "3xtf32"])
    ap.add_argument("--workload", default="sun_train", choices=["sun_train", "inference", "sky", "trunk", "trunk_train"],
                    help="sun_train: the sun-position pre-train step, BASELINE configs[1] (default; the line also carries the full-inference "
                         "throughput, configs[0]); inference: full generator inference; sky: its sky branch; trunk: the DA residual trunk alone")
    ap.add_argument("--no-inference", action="store_true", help="sun_train: skip the secondary full-inference measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
