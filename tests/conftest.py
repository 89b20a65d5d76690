import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "hdr-map-reconstruction-from-a-single-ldr-sky-panoramic-image-for-outdoor-illumination-estimation_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (directory name has hyphens, so it is imported by string)."""
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "da_golden.npz"))


def golden_cases(g):
    out = []
    for line in g["cases"]:
        name, kind, B, h, w, C, F, k, dil, sky, oh, ow = str(line).split("|")
        out.append(dict(name=name, kind=kind, B=int(B), h=int(h), w=int(w), C=int(C), F=int(F), k=int(k),
                        dilation=int(dil), skydome=bool(int(sky)), out_hw=(int(oh), int(ow))))
    return out
