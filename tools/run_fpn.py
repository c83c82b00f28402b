#!/usr/bin/env python
"""One view through the FPN hand-off at BASELINE configs[1] shapes (for ncu): 1x1 output convolutions of the library for
5 views per stage, then the cost volume from the prepared maps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import mdf_net_b200 as mdf
from mdf_net_b200 import ops, synthetic as syn

cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
K, E = syn.camera_rig(1, 5, 1152, 1600, seed=1)
with torch.no_grad():
    for rep in range(2):
        for s in range(3):
            H, W = syn.stage_shapes(1152, 1600)[s]
            C, D, G = syn.STAGE_CHANNELS[s], syn.STAGE_DEPTHS[s], syn.STAGE_GROUPS[s]
            P = syn.projection_matrices(K, E, 2.0 ** (3 - s))
            gen = torch.Generator(device="cuda").manual_seed(70 + s)
            xs = [torch.nn.functional.avg_pool2d(torch.randn((1, 64, H, W), device="cuda", generator=gen), 3, 1, 1) * 3.0 for _ in range(5)]
            wt = torch.randn((C, 64), device="cuda", generator=gen) * 0.08
            hyp = cu(syn.uniform_hypos(1, D) if s == 0 else syn.scene_hypos(1, D, H, W, seed=1))
            m = mdf.VectorAggregate(G).cuda().eval()
            cwt = m.depth_weight[0].conv.weight
            q4, cq4 = ops.fpn_out_prepped(xs[0], wt, G, cwt, True)
            s4 = torch.stack([ops.fpn_out_prepped(x, wt, G, cwt, False)[0] for x in xs[1:]], 0)
            m(mdf.PreppedFeatures(q4, cq4, s4), cu(P[:, 0]), [cu(P[:, v]) for v in range(1, 5)], hyp)
torch.cuda.synchronize()
print("done")
