#!/usr/bin/env python
"""Error statistics of the CUDA cost volume vs the CPU oracle (f32) and its float64 evaluation.
Run on a GPU box:  python tools/diag_parity.py [h0 w0 nviews]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

from oracle import c_oracle as co
from test_gpu_parity import run_cost_volume, stage_case


def stats(a, b):
    a = a.astype(np.float64); b = b.astype(np.float64)
    d = np.abs(a - b)
    return (f"rel_l2 {np.linalg.norm(a - b) / np.linalg.norm(b):.2e}  max_abs/max {d.max() / np.abs(b).max():.2e}  "
            f"frac>1e-5 {np.mean(d > 1e-5):.2e}  frac>1e-6 {np.mean(d > 1e-6):.2e}")


h0, w0, n = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (512, 640, 3)
for stage in range(3):
    c = stage_case(stage, h0, w0, n)
    kw = dict(ref_proj=c["ref_proj"], src_projs=c["src_projs"])
    f32 = co.vector_aggregate(c["features"], c["hypos"], c["params"], c["G"], **kw)
    f64 = co.vector_aggregate(c["features"], c["hypos"], c["params"], c["G"], prec="f64", **kw)
    print(f"stage {stage}: oracle f32 vs f64   {stats(f32, f64)}")
    for algo in (1, 2):
        out = run_cost_volume(c["features"], c["ref_proj"], c["src_projs"], c["hypos"], c["params"], c["G"], algo)
        print(f"  cuda algo {algo} vs oracle f32 {stats(out, f32)}")
        print(f"  cuda algo {algo} vs f64        {stats(out, f64)}")
