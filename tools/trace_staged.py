#!/usr/bin/env python
"""Phase timestamps of the staged hot kernel (algo 31 = tracing variant): where a CTA's lifetime goes.
    python tools/trace_staged.py [workload]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

os.environ.setdefault("MDF_B200_TUNING", "1")      # the variants live in the tuning build
import bench
from mdf_net_b200 import _cabi, ops

workload = sys.argv[1] if len(sys.argv) > 1 else "dtu_1600x1152_n5"
h0, w0, nviews, batch = bench.WORKLOADS[workload]
view = bench.make_view(h0, w0, nviews, batch, seed=1)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
lib = _cabi.lib()
W = 32
buf = np.zeros((4096, W), np.int64)
names = ["start", "pre-barrier", "barrier", "pos0", "announce0", "tiles"]
for s, st in enumerate(view):
    p = st["params"]
    f32 = lambda v: cu(np.asarray(v, np.float32).reshape(-1))
    args = ([cu(f) for f in st["features"]], cu(st["ref_proj"]), [cu(q) for q in st["src_projs"]], cu(st["hypos"]),
            f32(p["cw"]), f32(p["bn_weight"]), f32(p["bn_bias"]), f32(p["bn_mean"]), f32(p["bn_var"]), float(p["bn_eps"]),
            f32(p["fc_weight"]), f32(p["fc_bias"]), st["G"])
    for _ in range(2):
        ops.cost_volume(*args, 31)
        torch.cuda.synchronize()
        n = lib.mdf_debug_read_trace(buf.ctypes.data_as(ctypes.c_void_p), 4096)
    t = buf[:n].copy()
    cnt = int(t[0, W - 2])
    d = np.diff(t[:, :cnt], axis=1).astype(np.float64)
    V = nviews - 1
    labels = ["hypos+rt issued", "barrier", "positions v0 (hypotheses arrive)", "announce v0", "wait tiles"]
    for v in range(V):
        labels += [f"prepare v{v + 1}", f"wait box v{v}", f"gather v{v}"]
    labels += ["early epilogue", "tail (vote, retries, late epilogue)"]
    life = (t[:, cnt - 1] - t[:, 0]).astype(np.float64)
    print(f"stage {s} G{st['G']}: {n} traced warps, lifetime mean {life.mean():.0f} cycles (p10 {np.percentile(life, 10):.0f}, p90 {np.percentile(life, 90):.0f})")
    for k in range(cnt - 1):
        print(f"    {labels[k] if k < len(labels) else k:40s} mean {d[:, k].mean():8.0f}  p50 {np.percentile(d[:, k], 50):8.0f}  p90 {np.percentile(d[:, k], 90):8.0f}   {100 * d[:, k].mean() / life.mean():5.1f} %")
