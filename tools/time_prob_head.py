#!/usr/bin/env python
"""The regulariser tail at 1600x1152: this repo's single fused launch (prob conv + softmax + depth
regression + confidence / curve fit) vs what the reference's chain costs on the same GPU -- cuDNN Conv3d(c0,1,3) + ATen
softmax / regression (+ this repo's unfused head) -- and the split path of this repo (cuDNN conv -> softmax_regress_fit)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import torch.nn.functional as F
from mdf_net_b200 import ops, synthetic as syn


def timeit(fn, n=9, warm=3, reps=10):
    """Median over n replays of a CUDA graph holding `reps` calls (the Python / ctypes cost of a call, ~40 us, is not
    what is being measured), in us per call."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    keep = []
    with torch.cuda.graph(g):
        for _ in range(reps):
            keep.append(fn())
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / reps)
    return sorted(ts)[len(ts) // 2] * 1e3


cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
tot = [0.0, 0.0, 0.0, 0.0]
for tf32 in (False,):
    torch.backends.cudnn.allow_tf32 = tf32
    for stage in range(3):
        H, W = syn.stage_shapes(1152, 1600)[stage]
        D, C = syn.STAGE_DEPTHS[stage], (16, 8, 8)[stage]
        g = torch.Generator(device="cuda").manual_seed(stage)
        x = torch.randn((1, C, D, H, W), device="cuda", generator=g).relu_()
        w = torch.randn((1, C, 3, 3, 3), device="cuda", generator=g) * 0.35
        hyp = cu(syn.uniform_hypos(1, D)) if stage == 0 else cu(syn.scene_hypos(1, D, H, W, seed=4))
        curve = ("gauss1", "laplace", "")[stage]
        last = stage == 2

        def fused():
            return ops.prob_head(x, w, hyp, curve, want_logits=False, want_prob=False, want_confidence=last)

        def fused_with_prob():
            return ops.prob_head(x, w, hyp, curve, want_logits=False, want_prob=True, want_confidence=last)

        def split():
            lg = F.conv3d(x, w, padding=1).squeeze(1)
            if curve:
                return ops.softmax_regress_fit(lg, hyp, curve, want_prob=False, want_confidence=False)
            return ops.softmax_regress(lg, hyp, False, True, 4, 1, 2, 2)

        def aten():
            p = F.softmax(F.conv3d(x, w, padding=1).squeeze(1), dim=1)
            return p, (p * hyp).sum(1)

        def conv_only():
            return F.conv3d(x, w, padding=1)

        for algo in (1, 2):
            ta = timeit(lambda: ops.prob_head(x, w, hyp, curve, want_logits=False, want_prob=False, want_confidence=last, algo=algo))
            print(f"   algo {algo}: {ta:.1f} us")
        t = [timeit(fused), timeit(fused_with_prob), timeit(split), timeit(aten), timeit(conv_only)]
        mb = x.numel() * 4 / 1e6
        print(f"stage {stage} C{C} D{D} {H}x{W}: fused {t[0]:.1f} us ({mb / t[0] * 1e3 / 1e3:.0f} GB/s of x, "
              f"{27 * C * D * H * W / t[0] / 1e6:.2f} TFMA/s)  +prob {t[1]:.1f} us | cuDNN conv + fused head {t[2]:.1f} us | "
              f"cuDNN conv + ATen softmax/regress {t[3]:.1f} us | cuDNN conv alone {t[4]:.1f} us")
        for i in range(4):
            tot[i] += t[i]
print(f"sum over the stages: fused {tot[0]:.1f} us, fused+prob {tot[1]:.1f} us, cuDNN conv + fused head {tot[2]:.1f} us, "
      f"cuDNN + ATen {tot[3]:.1f} us")
