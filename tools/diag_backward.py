#!/usr/bin/env python
"""Backward error statistics: CUDA vs float64 autograd of the restatement, next to float32 autograd of the same."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from test_gpu_backward import case, make_module, run_module, reference_grads
from conftest import rel_l2
for stage in (0, 1, 2):
    for training in (False, True):
        feats, ref_proj, src_projs, hyp, p, G, gout = case(stage, 256, 320, 4, 2, seed=500)
        m = make_module(G, p); m.train(training)
        r = run_module(m, feats, ref_proj, src_projs, hyp, gout)
        r64 = reference_grads(feats, ref_proj, src_projs, hyp, p, G, gout, training, torch.float64)
        r32 = reference_grads(feats, ref_proj, src_projs, hyp, p, G, gout, training, torch.float32)
        print(f"stage {stage} train={training}: gf cuda-vs-f64 {rel_l2(r['gf'], r64['gf']):.2e}  f32-vs-f64 {rel_l2(r32['gf'], r64['gf']):.2e}  "
              f"cuda-vs-f32 {rel_l2(r['gf'], r32['gf']):.2e} | gbn cuda {r['gbn']} f32 {r32['gbn']} f64 {r64['gbn']} | gcw {rel_l2(r['gcw'], r64['gcw']):.1e}/{rel_l2(r32['gcw'], r64['gcw']):.1e}")
