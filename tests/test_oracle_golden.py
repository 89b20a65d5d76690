"""The oracle against the golden vectors produced by the reference's own source (tests/golden/make_golden.py).
Offsets, gather indices and bilinear weights: bit for bit.  Layer outputs: tolerance (matmul order unspecified)."""
import numpy as np
import pytest
import torch

from conftest import golden_cases
from oracle import da_oracle as O


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.int32)


def test_offset_tables_bit_exact(golden):
    n = 0
    for line in golden["off_cases"]:
        name, h, w, k, dil, sky = str(line).split("|")
        want = golden[f"offonly__{name}"]
        got = O.offsets(int(h), int(w), int(k), int(dil), bool(int(sky)))
        assert np.array_equal(_bits(got), _bits(want)), name     # NaNs (k=7 at 8x32) compare by bit pattern too
        n += 1
    assert n == len(golden["off_cases"]) and n >= 17


def test_layer_offsets_indices_weights_bit_exact(golden):
    for c in golden_cases(golden):
        name = c["name"]
        h, w = (c["h"], c["w"]) if c["kind"] == "conv" else c["out_hw"]
        off = O.offsets(h, w, c["k"], c["dilation"], c["skydome"])
        assert np.array_equal(_bits(off), _bits(golden[f"off__{name}"])), name
        for s in (O.sample(h, w, c["k"], off), O.sample_np(h, w, c["k"], off)):
            idx = golden[f"idx__{name}"]          # [4, h, w, k2, 2] (y, x): corners (y0,x0) (y0,x1) (y1,x0) (y1,x1)
            for ci, (yy, xx) in enumerate((("y0", "x0"), ("y0", "x1"), ("y1", "x0"), ("y1", "x1"))):
                assert np.array_equal(s[yy], idx[ci, ..., 0]), (name, ci, "y")
                assert np.array_equal(s[xx], idx[ci, ..., 1]), (name, ci, "x")
            wg = golden[f"wgt__{name}"]
            for ci in range(4):
                assert np.array_equal(_bits(s[f"w{ci}"]), _bits(wg[ci])), (name, ci, "w")


def test_layer_outputs_match_reference_source(golden):
    for c in golden_cases(golden):
        name = c["name"]
        x, kern, b, want = (golden[f"{k}__{name}"] for k in ("x", "k", "b", "y"))
        if c["kind"] == "conv":
            got = O.conv2d_forward(x, kern, b, c["k"], c["dilation"], c["skydome"])
        else:
            r = O.resize_bilinear(x, *c["out_hw"])
            assert np.array_equal(_bits(r.numpy()), _bits(golden[f"resized__{name}"])), name
            got = O.deconv2d_forward(x, kern, b, c["out_hw"], c["k"], c["dilation"], c["skydome"])
        got = got.numpy()
        rel = np.linalg.norm(got - want) / np.linalg.norm(want)
        assert rel < 2e-6, (name, rel)


def test_reference_failure_modes(golden):
    assert int(golden["undefined_2x8_k3"]) == 1
    with pytest.raises(Exception, match="undefined coordinates"):
        O.offsets(2, 8, 3)
    assert int(golden["even_k_asserts"]) == 1
    with pytest.raises(AssertionError):
        O.offsets(8, 32, 4)
    assert int(golden["k1_raises"]) == 1


def test_gather_indices_in_range_all_configs():
    # SURVEY notes: after the reference's two wraps every index is inside the padded map for every config size
    for (h, w) in [(8, 32), (16, 64), (32, 128), (64, 256)]:
        s = O.sample(h, w, 3, O.offsets(h, w, 3))
        assert s["rc"] == 0
        assert s["x0"].min() >= 0 and s["x1"].max() <= w + 1
    s = O.sample(32, 128, 7, O.offsets(32, 128, 7))
    assert s["rc"] == 0


def test_dead_bottom_row_taps():
    # when y clips to in_h-1 both y weights vanish -> 3*w dead samples in the last row (k=3)
    h, w, k = 8, 32, 3
    s = O.sample(h, w, k, O.offsets(h, w, k))
    dead = (s["w0"] == 0) & (s["w1"] == 0) & (s["w2"] == 0) & (s["w3"] == 0)
    assert dead[h - 1].sum() == 3 * w
    assert dead[: h - 1].sum() == 0


def test_backward_matches_fp64_autograd():
    rng = np.random.default_rng(3)
    B, h, w, C, F, k = 1, 4, 16, 3, 2, 3
    x = rng.standard_normal((B, h, w, C)).astype(np.float32)
    kern = rng.standard_normal((k * k * C, F)).astype(np.float32)
    b = rng.standard_normal(F).astype(np.float32)
    dy = rng.standard_normal((B, h, w, F)).astype(np.float32)
    g32 = O.conv2d_backward(x, kern, b, dy, k)
    g64 = O.conv2d_backward(x, kern, b, dy, k, acc_dtype=torch.float64)
    for a, c in zip(g32, g64):
        assert torch.allclose(a.double(), c.double(), rtol=1e-4, atol=1e-5)
