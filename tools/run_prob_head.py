#!/usr/bin/env python
"""Profiling target: the fused regulariser tail of the three stages at 1600x1152, `reps` times (default 2).
ncu --set full -k regex:prob_head_kernel -s 3 -c 3 python tools/run_prob_head.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mdf_net_b200 import ops, synthetic as syn

cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
algo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ins = []
for stage in range(3):
    H, W = syn.stage_shapes(1152, 1600)[stage]
    D, C = syn.STAGE_DEPTHS[stage], (16, 8, 8)[stage]
    g = torch.Generator(device="cuda").manual_seed(stage)
    x = torch.randn((1, C, D, H, W), device="cuda", generator=g).relu_()
    w = torch.randn((1, C, 3, 3, 3), device="cuda", generator=g) * 0.35
    hyp = cu(syn.uniform_hypos(1, D)) if stage == 0 else cu(syn.scene_hypos(1, D, H, W, seed=4))
    ins.append((x, w, hyp, ("gauss1", "laplace", "")[stage], stage == 2))
for _ in range(reps):
    for x, w, hyp, curve, last in ins:
        ops.prob_head(x, w, hyp, curve, want_logits=False, want_prob=False, want_confidence=last, algo=algo)
torch.cuda.synchronize()
print("ok")
