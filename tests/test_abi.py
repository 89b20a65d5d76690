"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol the header declares, the
host-side entry point (offset table, libm) is bit-exact with the oracle, and error behaviour mirrors the reference."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from oracle import da_oracle as O


def test_header_symbols_exported_and_bound(pkg):
    hdr = open(os.path.join(ROOT, "include", "skydome_b200.h")).read()
    declared = set(re.findall(r"\b(sky_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 10
    assert declared == set(pkg._lib.SIGNATURES), declared ^ set(pkg._lib.SIGNATURES)
    for name in declared:
        assert hasattr(pkg._lib.LIB, name), name
    assert pkg._lib.LIB.sky_version() >= 100


def _host_offsets(pkg, h, w, k, dil=1, sky=True):
    out = np.empty((h, k * k, 2), np.float32)
    pkg._lib.check(pkg._lib.LIB.sky_da_offsets_host(h, w, k, dil, int(sky), out.ctypes.data))
    return out


def test_host_offsets_bit_exact_vs_oracle_and_golden(pkg, golden):
    for line in golden["off_cases"]:
        name, h, w, k, dil, sky = str(line).split("|")
        want = golden[f"offonly__{name}"]
        if np.isnan(want).any():
            with pytest.raises(pkg._lib.SkydomeError):
                _host_offsets(pkg, int(h), int(w), int(k), int(dil), bool(int(sky)))
            continue
        got = _host_offsets(pkg, int(h), int(w), int(k), int(dil), bool(int(sky)))
        assert np.array_equal(got.view(np.int32), want.view(np.int32)), name
    for (h, w, k, dil, sky) in [(24, 96, 3, 1, True), (48, 192, 5, 1, False), (32, 128, 3, 2, True), (33, 131, 3, 1, True)]:
        got = _host_offsets(pkg, h, w, k, dil, sky)
        assert np.array_equal(got.view(np.int32), O.offsets(h, w, k, dil, sky).view(np.int32)), (h, w, k, dil, sky)


def test_error_behaviour_mirrors_reference(pkg):
    with pytest.raises(AssertionError, match="kernel_size must be odd"):     # distortion_aware_ops.py:188
        _host_offsets(pkg, 8, 32, 4)
    with pytest.raises(Exception, match="undefined coordinates"):            # :252
        _host_offsets(pkg, 2, 8, 3)
    with pytest.raises(ValueError):
        _host_offsets(pkg, 0, 32, 3)
    with pytest.raises(pkg._lib.SkydomeError):                               # k=1: the reference cannot build it either
        _host_offsets(pkg, 8, 32, 1)


def test_layer_signatures_match_reference(pkg):
    import inspect
    c = inspect.signature(pkg.conv2d.__init__)
    assert list(c.parameters)[:9] == ["self", "filters", "kernel_size", "strides", "padding", "dilation_rate",
                                      "kernel_initializer", "bias_initializer", "skydome"]
    assert c.parameters["padding"].default == "VAILD" and c.parameters["kernel_size"].default == 3
    d = inspect.signature(pkg.deconv2d.__init__)
    assert list(d.parameters)[:10] == ["self", "filters", "kernel_size", "strides", "output_imshape", "padding",
                                       "dilation_rate", "skydome", "kernel_initializer", "bias_initializer"]


def test_no_cpu_path(pkg):
    import torch
    layer = pkg.conv2d(4, device="cpu") if False else None   # building needs CUDA memory; the call must refuse CPU tensors
    with pytest.raises(RuntimeError, match="no CPU path"):
        pkg.distortion_aware_ops._require_cuda(torch.zeros(1, 4, 16, 4), "inputs")
    assert layer is None
