#!/usr/bin/env python
"""HyposByFit at 1600x1152: this repo's two kernels vs the same formulae in plain PyTorch (ATen eager) on the same GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import torch.nn.functional as F
import mdf_net_b200 as mdf
from mdf_net_b200 import synthetic as syn


def timeit(fn, n=7, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def aten_laplace(depth, prob, hyp, dr, nd, thresh):
    y = torch.log(prob.clamp(min=1e-40)); x = (hyp - depth.unsqueeze(1)).abs()
    s = 1 / ((x * y).sum(1) / (x * x).sum(1)).abs()
    s = F.interpolate(s.unsqueeze(1), scale_factor=2, mode="bilinear").squeeze(1)
    d = F.interpolate(depth.unsqueeze(1), scale_factor=2, mode="bilinear").squeeze(1)
    res = (s * np.log(thresh)).abs().clamp(min=1e-6, max=float(dr[0, 1] - dr[0, 0]) * 0.2)
    k = torch.arange(nd, device=d.device, dtype=d.dtype).view(1, nd, 1, 1)
    return ((d - 0.5 * res).unsqueeze(1) + (res / (nd - 1)).unsqueeze(1) * k).clamp(float(dr[0, 0]), float(dr[0, 1]))


cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
dr = cu(np.array([[425.0, 935.0]], np.float32))
for stage, (curve, thresh) in enumerate((("gauss1", 0.95), ("laplace", 1e-5))):
    H, W = syn.stage_shapes(1152, 1600)[stage]
    D, ND = syn.STAGE_DEPTHS[stage], syn.STAGE_DEPTHS[stage + 1]
    prob = torch.softmax(cu(syn.regulariser_logits(1, D, H, W, seed=3, peak=6.0)), 1)
    hyp = cu(syn.uniform_hypos(1, D)) if stage == 0 else cu(syn.scene_hypos(1, D, H, W, seed=4))
    depth = (prob * hyp).sum(1)
    m = mdf.HyposByFit(ND, curve, thresh)
    t = timeit(lambda: m(depth, dr, prob, hyp, upsample=True))
    line = f"stage {stage}->{stage + 1} {curve}: fused {t * 1e3:.1f} us"
    if curve == "laplace":
        line += f"   ATen eager {timeit(lambda: aten_laplace(depth, prob, hyp, dr, ND, thresh)) * 1e3:.1f} us"
    print(line)
