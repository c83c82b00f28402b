#!/usr/bin/env python
"""Run the fused cost volume of the three stages of one view (for ncu):  python tools/run_cv.py [algo] [workload] [scene|wide] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

os.environ.setdefault("MDF_B200_TUNING", "1")      # the variants live in the tuning build
import bench
from mdf_net_b200 import ops, synthetic as syn

algo = int(sys.argv[1]) if len(sys.argv) > 1 else 0
workload = sys.argv[2] if len(sys.argv) > 2 else "dtu_1600x1152_n5"
kind = sys.argv[3] if len(sys.argv) > 3 else "scene"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
h0, w0, nviews, batch = bench.WORKLOADS[workload]
view = bench.make_view(h0, w0, nviews, batch, seed=1)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
calls = []
for s, st in enumerate(view):
    if s > 0 and kind == "wide":
        st["hypos"] = syn.scene_hypos(batch, st["D"], st["H"], st["W"], seed=1, range_mm=(40.0, 102.0))
    p = st["params"]
    f32 = lambda v: cu(np.asarray(v, np.float32).reshape(-1))
    calls.append(([cu(f) for f in st["features"]], cu(st["ref_proj"]), [cu(q) for q in st["src_projs"]], cu(st["hypos"]),
                  f32(p["cw"]), f32(p["bn_weight"]), f32(p["bn_bias"]), f32(p["bn_mean"]), f32(p["bn_var"]), float(p["bn_eps"]),
                  f32(p["fc_weight"]), f32(p["fc_bias"]), st["G"]))
from mdf_net_b200 import _cabi
for _ in range(reps):
    for c in calls:
        try:
            ops.cost_volume(*c, algo)
        except _cabi.MdfError:
            pass
torch.cuda.synchronize()
print("done")
