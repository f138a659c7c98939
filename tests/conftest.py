"""pytest configuration: the `gpu` marker, import paths, shared helpers.

* `-m "not gpu"`: oracle vs the golden fixtures generated from the reference, host logic, C-ABI exports.
* `-m gpu`      : parity tests proper — the CUDA path (through the C ABI) vs oracle / golden fixtures.
Nothing here reads /root/reference (it does not exist on the GPU box); the one test that does
(tests/test_oracle_vs_reference.py) skips itself when the reference is not mounted.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lrp-imagecaptioning-pytorch_b200")
for p in (PKG, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    out = {}
    for k in d.files:
        v = d[k]
        out[k] = torch.from_numpy(v) if v.dtype.kind in "fiub" and v.ndim > 0 else v
    return out


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]
    return get


def assert_close(actual, expected, rtol=1e-4, atol=1e-6, what=""):
    """north_star fp32 bar: rtol 1e-4 / atol 1e-6 (elementwise, like torch.allclose)."""
    actual = actual.detach().cpu().double()
    expected = expected.detach().cpu().double()
    assert actual.shape == expected.shape, f"{what}: shape {tuple(actual.shape)} != {tuple(expected.shape)}"
    err = (actual - expected).abs()
    bound = atol + rtol * expected.abs()
    bad = err > bound
    if bad.any():
        i = int(torch.argmax(err - bound))
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.numel()} elements out of tolerance; worst "
                             f"actual={actual.flatten()[i].item():.8g} expected={expected.flatten()[i].item():.8g} "
                             f"(max |expected|={expected.abs().max().item():.4g})")


def spearman(a, b):
    a = a.detach().cpu().double().flatten()
    b = b.detach().cpu().double().flatten()
    ra = torch.empty_like(a); ra[a.argsort()] = torch.arange(a.numel(), dtype=torch.double)
    rb = torch.empty_like(b); rb[b.argsort()] = torch.arange(b.numel(), dtype=torch.double)
    ra -= ra.mean(); rb -= rb.mean()
    return float((ra * rb).sum() / (ra.norm() * rb.norm()))
