#!/usr/bin/env python
"""homo_warping (row a1) and homo_aggregate_by_variance (row a4) at BASELINE configs[1] shapes: time and HBM rate
(the output volume (B,C,D,H,W) written once is the traffic that bounds both)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mdf_net_b200 import ops, synthetic as syn


def timeit(fn, n=7, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2] * 1e3


cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
h0, w0, N = 1152, 1600, 5
K, E = syn.camera_rig(1, N, h0, w0, seed=1)
for s in range(3):
    H, W = syn.stage_shapes(h0, w0)[s]
    C, D = syn.STAGE_CHANNELS[s], syn.STAGE_DEPTHS[s]
    P = syn.projection_matrices(K, E, 2.0 ** (3 - s))
    feats = [cu(f) for f in syn.smooth_features(1, N, C, H, W, seed=10 + s)]
    hyp = cu(syn.uniform_hypos(1, D) if s == 0 else syn.scene_hypos(1, D, H, W, seed=1))
    rp, sps = cu(P[:, 0]), [cu(P[:, v]) for v in range(1, N)]
    out_b = C * D * H * W * 4
    t_w = timeit(lambda: ops.homo_warp(feats[1], sps[0], rp, hyp))
    t_v = timeit(lambda: ops.variance_volume(feats, rp, sps, hyp))
    print(f"stage {s} C{C} D{D} {H}x{W}: homo_warp {t_w:.1f} us ({out_b / 1e9 / (t_w / 1e6):.0f} GB/s of output), "
          f"variance_volume {t_v:.1f} us ({out_b / 1e9 / (t_v / 1e6):.0f} GB/s of output); output {out_b / 1e6:.0f} MB")
