#!/usr/bin/env python
"""Time every tuning variant of the staged cost-volume kernel (op level: setup + prep + hot kernel) on
BASELINE.json's config-2 shapes.  Run on a GPU box:  python tools/tune_staged.py [rough]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

os.environ.setdefault("MDF_B200_TUNING", "1")      # the variants live in the tuning build
import bench
from mdf_net_b200 import _cabi, ops, synthetic as syn

rough = "rough" in sys.argv
h0, w0, nviews, batch = bench.WORKLOADS["dtu_1600x1152_n5"]
view = bench.make_view(h0, w0, nviews, batch, seed=1)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
cvb, _ = bench.algorithmic_bytes(h0, w0, nviews, batch)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for s, st in enumerate(view):
    if rough and s > 0:
        st["hypos"] = syn.pixel_hypos(batch, st["D"], st["H"], st["W"], seed=5, smooth="iid" not in sys.argv)
    if "wide" in sys.argv and s > 0:     # the widest search range HyposByFit allows: 0.2 * (dmax - dmin) = 102 mm
        st["hypos"] = syn.scene_hypos(batch, st["D"], st["H"], st["W"], seed=1, range_mm=(40.0, 102.0))
    p = st["params"]
    f32 = lambda v: cu(np.asarray(v, np.float32).reshape(-1))
    args = ([cu(f) for f in st["features"]], cu(st["ref_proj"]), [cu(q) for q in st["src_projs"]], cu(st["hypos"]),
            f32(p["cw"]), f32(p["bn_weight"]), f32(p["bn_bias"]), f32(p["bn_mean"]), f32(p["bn_var"]), float(p["bn_eps"]),
            f32(p["fc_weight"]), f32(p["fc_bias"]), st["G"])
    base = None
    for variant in range(8):
        try:
            out = ops.cost_volume(*args, 16 + variant)
        except _cabi.MdfError:
            break
        torch.cuda.synchronize()
        if base is None:
            base = out.clone()
        err = float((out - base).abs().max())
        ts = []
        for _ in range(9):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.cost_volume(*args, 16 + variant); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        print(f"stage {s} G{st['G']} variant {variant}: {t * 1e3:8.1f} us  {cvb[s] / 1e9 / (t / 1e3):7.0f} GB/s  maxdiff vs v0 {err:.1e}", flush=True)
