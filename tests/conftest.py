"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"`: oracle vs golden vectors, host logic, C-ABI export check (no GPU needed).
`-m gpu`      : parity of the CUDA path (through the C-ABI) against the oracle / goldens.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """The CUDA library and the oracle are built artefacts (git-ignored): build whatever is missing.  An existing
    library is used as it is (the GPU box runs the files that travelled with the snapshot)."""
    try:
        from mdf_net_b200 import build
        if not os.path.exists(build.LIB_PATH):
            build.build_library()
    except Exception as e:  # pragma: no cover - reported by the tests that need the library
        print(f"[conftest] could not build libmdf_b200.so: {e}")
    try:
        from oracle import c_oracle
        c_oracle.build()
    except Exception as e:  # pragma: no cover
        print(f"[conftest] could not build the oracle: {e}")
    try:
        from oracle import ref_install
        ref_install.install()           # no-op where /root/reference does not exist (the GPU box uses the shipped copy)
    except Exception as e:  # pragma: no cover
        print(f"[conftest] could not install oracle/_ref: {e}")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def max_abs_over_max(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
