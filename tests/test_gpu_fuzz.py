"""A short randomised sweep of shapes through the CUDA paths (tools/fuzz_gpu.py): staged = direct = oracle for the eval forward,
train-mode forward + backward against float64 autograd, variance volume against the oracle.  Fixed seed; the long runs are
`python tools/fuzz_gpu.py 240 <seed>`."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_random_shapes_for_twenty_seconds():
    import fuzz_gpu
    r = fuzz_gpu.run(budget=20.0, seed=31)
    assert r["cases"]["forward"] >= 20 and r["cases"]["backward"] >= 5 and r["cases"]["variance"] >= 5, r
    assert r["worst"]["forward_vs_oracle"] < 5e-5 and r["worst"]["variance"] < 1e-5, r
