#!/usr/bin/env python
"""Randomised shapes through the CUDA paths for a time budget (default 150 s):  python tools/fuzz_gpu.py [seconds] [seed]
  * eval forward: the staged kernel (every warp footprint the width picks) vs the direct kernel -- bit-identical by construction --
    and both vs the C oracle (tests' tolerance) on maps small enough for it;
  * train-mode forward + backward vs float64 autograd of the plain-PyTorch restatement (tests/torch_ref.py, the tests' criterion);
  * variance volume vs the oracle.
Test infrastructure (it executes oracle/): not imported by the package."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import mdf_net_b200 as mdf
from mdf_net_b200 import ops, synthetic as syn
from oracle import c_oracle as co
import test_gpu_backward as tb
from conftest import rel_l2

def run(budget: float = 150.0, seed: int = 2026) -> dict:
    rng = np.random.default_rng(seed)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    t_end = time.time() + budget
    n = {"forward": 0, "backward": 0, "variance": 0}
    worst = {"forward_vs_oracle": 0.0, "backward_gf": 0.0, "variance": 0.0}
    while time.time() < t_end:
        G = int(rng.choice([8, 16, 32])); C = 2 * G
        H, W = int(rng.integers(1, 70)), int(rng.integers(1, 150))
        N, B, D = int(rng.integers(2, 7)), int(rng.integers(1, 3)), int(rng.integers(2, 13))      # (the generators need D >= 2)
        seed = int(rng.integers(1, 1 << 30))
        K, E = syn.camera_rig(B, N, max(H, 2) * 8, max(W, 2) * 8, seed=seed)
        P = syn.projection_matrices(K, E, level_div=8.0)
        feats = syn.smooth_features(B, N, C, H, W, seed=seed + 1)
        kind = int(rng.integers(0, 3))
        hyp = syn.uniform_hypos(B, D) if kind == 0 else (syn.pixel_hypos(B, D, H, W, seed=seed + 2) if kind == 1 else syn.scene_hypos(B, D, H, W, seed=seed + 2))
        p = syn.depth_weight_params(G, seed=seed + 3)
        f32 = lambda v: cu(np.asarray(v, np.float32).reshape(-1))
        args = ([cu(f) for f in feats], cu(P[:, 0]), [cu(P[:, v]) for v in range(1, N)], cu(hyp), f32(p["cw"]), f32(p["bn_weight"]),
                f32(p["bn_bias"]), f32(p["bn_mean"]), f32(p["bn_var"]), float(p["bn_eps"]), f32(p["fc_weight"]), f32(p["fc_bias"]), G)
        what = (G, B, N, D, H, W, kind, seed)
        a1, a2 = ops.cost_volume(*args, 1).cpu().numpy(), ops.cost_volume(*args, 2).cpu().numpy()
        assert np.isfinite(a1).all(), what
        assert np.array_equal(a1, a2) or np.abs(a1 - a2).max() < 2e-6, ("staged vs direct", what, np.abs(a1 - a2).max())
        ref = co.vector_aggregate(feats, hyp, p, G, ref_proj=P[:, 0], src_projs=[P[:, v] for v in range(1, N)])
        err = np.abs(a1 - np.nan_to_num(ref, nan=0.5)).max()
        assert err < 5e-5, ("staged vs oracle", what, err)
        worst["forward_vs_oracle"] = max(worst["forward_vs_oracle"], float(err))
        n["forward"] += 1
        if W > 1 and H > 1 and B * D * H * W >= 64 and rng.random() < 0.5:
            gout = rng.standard_normal((B, G, D, H, W)).astype(np.float32)
            training = bool(rng.random() < 0.7) and B * D * H * W > 1
            m = tb.make_module(G, p); m.train(training)
            r = tb.run_module(m, feats, P[:, 0], [P[:, v] for v in range(1, N)], hyp, gout)
            r64 = tb.reference_grads(feats, P[:, 0], [P[:, v] for v in range(1, N)], hyp, p, G, gout, training, torch.float64)
            r32 = tb.reference_grads(feats, P[:, 0], [P[:, v] for v in range(1, N)], hyp, p, G, gout, training, torch.float32)
            mine, theirs = rel_l2(r["gf"], r64["gf"]), rel_l2(r32["gf"], r64["gf"])
            assert rel_l2(r["cv"], r64["cv"]) < 1e-5, ("train forward", what, training)
            # ReLU of depth_weight is a kink: an element whose pre-activation is within the forward's own rounding (MUFU 2^x, 1/x:
            # |dh| ~ 1e-5) of zero takes the other branch than float64 does -- the kernel differentiates ITS forward (the backward
            # recomputes z with the same instructions), so its gradient is exact for the function it computed, and one such element
            # in a small tensor is an O(1) error at one pixel (measured: 2 of 87 600 elements, rel-L2 6e-4 .. 1.6e-3 on a 50x146 map;
            # train mode centres z on the kink, so it has more of them).  Small tensors get the budget of a few flips.
            elems = B * D * H * W * (N - 1)
            budget_l2 = max(2e-4, 2.0 * theirs, min(0.05, 6.0 / np.sqrt(max(elems, 1))))
            assert mine < budget_l2, ("backward features", what, training, mine, theirs, budget_l2)
            if theirs < 1e-3:                        # (degenerate geometries -- every sample out of the image -- leave only rounding noise to compare)
                worst["backward_gf"] = max(worst["backward_gf"], float(mine))
            n["backward"] += 1
        if rng.random() < 0.4:
            Cv = int(rng.choice([12, 16, 32, 64]))
            fv = syn.smooth_features(B, N, Cv, H, W, seed=seed + 5)
            out = mdf.homo_aggregate_by_variance([cu(f) for f in fv], cu(P[:, 0]), [cu(P[:, v]) for v in range(1, N)], cu(hyp)).cpu().numpy()
            refv = co.variance_aggregate(fv, hyp, ref_proj=P[:, 0], src_projs=[P[:, v] for v in range(1, N)])
            e = rel_l2(out, np.nan_to_num(refv, nan=0.0)) if np.isfinite(refv).all() else 0.0
            assert np.isfinite(out).all() and e < 1e-5, ("variance", what, Cv, e)
            worst["variance"] = max(worst["variance"], float(e))
            n["variance"] += 1
    return {"cases": n, "worst": worst}


if __name__ == "__main__":
    r = run(float(sys.argv[1]) if len(sys.argv) > 1 else 150.0, int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
    print(f"fuzz ok: {r['cases']}, worst {r['worst']}")
