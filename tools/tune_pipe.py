#!/usr/bin/env python
"""Time the pipelined cost-volume kernel (mdf_pipe.cuh) against the staged one, hot kernel only (CUDA events recorded
inside the library around the launch), and compare the outputs bit for bit.

    python tools/tune_pipe.py [workload] [scene|wide|rough|iid] [algos=16,32,33,...] [rg=0,1,2,3]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

os.environ.setdefault("MDF_B200_TUNING", "1")      # the variants live in the tuning build
import bench
from mdf_net_b200 import _cabi, ops, synthetic as syn

args = [a for a in sys.argv[1:] if "=" not in a]
kw = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
workload = next((a for a in args if a in bench.WORKLOADS), "dtu_1600x1152_n5")
kind = next((a for a in args if a in ("scene", "wide", "rough", "iid")), "scene")
algos = [int(x) for x in kw.get("algos", "16,32,33,34").split(",")]
rgs = [int(x) for x in kw.get("rg", "0").split(",")]
stages = [int(x) for x in kw.get("stages", "0,1,2").split(",")]
reps = int(kw.get("reps", "9"))

h0, w0, nviews, batch = bench.WORKLOADS[workload]
view = bench.make_view(h0, w0, nviews, batch, seed=1)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
cvb, _ = bench.algorithmic_bytes(h0, w0, nviews, batch)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
lib = _cabi.lib()
total = {}
for s, st in enumerate(view):
    if s not in stages:
        continue
    if s > 0 and kind == "wide":      # the widest search range HyposByFit allows: 0.2 * (dmax - dmin) = 102 mm
        st["hypos"] = syn.scene_hypos(batch, st["D"], st["H"], st["W"], seed=1, range_mm=(40.0, 102.0))
    if s > 0 and kind in ("rough", "iid"):
        st["hypos"] = syn.pixel_hypos(batch, st["D"], st["H"], st["W"], seed=5, smooth=kind == "rough")
    p = st["params"]
    f32 = lambda v: cu(np.asarray(v, np.float32).reshape(-1))
    cv_args = ([cu(f) for f in st["features"]], cu(st["ref_proj"]), [cu(q) for q in st["src_projs"]], cu(st["hypos"]),
               f32(p["cw"]), f32(p["bn_weight"]), f32(p["bn_bias"]), f32(p["bn_mean"]), f32(p["bn_var"]), float(p["bn_eps"]),
               f32(p["fc_weight"]), f32(p["fc_bias"]), st["G"])
    base = ops.cost_volume(*cv_args, 16).clone()
    torch.cuda.synchronize()
    for algo in algos:
        for rg in (rgs if algo >= 32 else [0]):
            code = algo + 256 * rg
            try:
                out = ops.cost_volume(*cv_args, code)
                torch.cuda.synchronize()
            except _cabi.MdfError as e:
                print(f"stage {s} algo {algo} rg {rg}: {e}", flush=True)
                continue
            diff = float((out - base).abs().max())
            nbad = int((out != base).sum())
            ts = []
            for _ in range(reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ops.time_next_hot_kernel(e0, e1)
                ops.cost_volume(*cv_args, code)
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = sorted(ts)[len(ts) // 2]
            total.setdefault((algo, rg), 0.0)
            total[(algo, rg)] += t
            print(f"{workload} {kind} stage {s} G{st['G']} algo {algo} rg {rg}: {t * 1e3:8.1f} us (min {min(ts) * 1e3:.1f})  "
                  f"{cvb[s] / 1e9 / (t / 1e3):7.0f} GB/s  maxdiff vs staged {diff:.1e} ({nbad} elements differ)", flush=True)
for k, t in total.items():
    print(f"sum over stages algo {k[0]} rg {k[1]}: {t * 1e3:.1f} us")
