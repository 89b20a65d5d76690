"""CPU checks of bench.py's host-side pieces: synthetic inputs and weights have the shapes / distributions SURVEY 8(d) states, the
reference arm runs, and the call counter of the C-ABI binding counts."""
import json
import subprocess
import sys

import numpy as np

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_synthetic_inputs():
    ldr = bench.make_ldr(4, 32, 128, seed=1)
    assert ldr.shape == (4, 32, 128, 3) and ldr.dtype == np.float32
    assert np.allclose(np.round(ldr * 255) / 255, ldr, atol=1e-7) and ldr.min() >= 0 and ldr.max() <= 1      # 8-bit codes in [0, 1]
    gt = bench.make_sunpose_gt(3, 32, 128, seed=2)
    assert gt.shape == (3, 4096) and np.allclose(gt.sum(1), 1, atol=1e-5) and (gt >= 0).all()
    assert (gt.max(1) > 10.0 / 4096).all()                      # a peaked von Mises-Fisher bump, not a flat map
    wg, ws = bench.make_inference_weights(32, 128)
    assert wg["conv1_u"][0].shape == (7, 7, 32, 3) and wg["sun"]["d4"]["kernel"].shape == (4, 4, 256, 512)
    assert ws["fc1"][0].shape == (8192, 4096) and ws["fc2"][0].shape == (4096, 4096) and ws["sunlayer1"]["conv1_kernel"].shape == (147, 32)
    lim = (6.0 / (8192 + 4096)) ** 0.5
    assert abs(np.abs(ws["fc1"][0]).max() - lim) < 1e-3 * lim   # glorot_uniform limit


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1", "--workload", "inference",
                          "--batch", "2"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stderr[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "panoramas/s" and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def test_abi_call_counter(pkg):
    lib = pkg._lib.LIB
    lib.counts = {}
    lib.sky_version()
    out = np.empty((8, 9, 2), np.float32)
    pkg._lib.check(lib.sky_da_offsets_host(8, 32, 3, 1, 1, out.ctypes.data))
    assert lib.counts == {"sky_version": 1, "sky_da_offsets_host": 1} and lib.launches() == 0
    lib.counts = None
