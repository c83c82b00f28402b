#!/usr/bin/env python
"""Where the fused regulariser tail spends its time: the convolution alone (logits out, no column tail) vs the whole launch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mdf_net_b200 import ops, _cabi, synthetic as syn

cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()


def timeit(fn, n=9, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / reps)
    return sorted(ts)[len(ts) // 2] * 1e3


lib = _cabi.lib()
for stage in range(3):
    H, W = syn.stage_shapes(1152, 1600)[stage]
    D, C = syn.STAGE_DEPTHS[stage], (16, 8, 8)[stage]
    g = torch.Generator(device="cuda").manual_seed(stage)
    x = torch.randn((1, C, D, H, W), device="cuda", generator=g).relu_()
    w = torch.randn((1, C, 3, 3, 3), device="cuda", generator=g) * 0.35
    hyp = cu(syn.uniform_hypos(1, D)) if stage == 0 else cu(syn.scene_hypos(1, D, H, W, seed=4))
    curve = ("gauss1", "laplace", "")[stage]
    logits = torch.empty((1, D, H, W), device="cuda")
    depth = torch.empty((1, H, W), device="cuda")
    s = torch.empty((1, H, W), device="cuda")
    conf = torch.empty((1, 2 * H, 2 * W), device="cuda")
    st = torch.cuda.current_stream
    pp = 0 if stage == 0 else 1

    def conv_only():
        r = lib.mdf_prob_head_fwd_ex(x.data_ptr(), w.data_ptr(), hyp.data_ptr(), pp, 1, C, D, H, W, logits.data_ptr(), None, None, None,
                                     4, 1, 2, 2, 0, None, 0, st().cuda_stream)
        assert r == 0, r

    def no_fit():
        r = lib.mdf_prob_head_fwd_ex(x.data_ptr(), w.data_ptr(), hyp.data_ptr(), pp, 1, C, D, H, W, None, None, depth.data_ptr(),
                                     conf.data_ptr() if stage == 2 else None, 4, 1, 2, 2, 0, None, 0, st().cuda_stream)
        assert r == 0, r

    def full():
        ops.prob_head(x, w, hyp, curve, want_logits=False, want_prob=False, want_confidence=stage == 2)

    if stage > 0:
        hu = cu(syn.uniform_hypos(1, D))

        def full_uniform():
            ops.prob_head(x, w, hu, curve, want_logits=False, want_prob=False, want_confidence=stage == 2)

        print(f"   stage {stage} with per-plane (uniform) hypotheses instead of per-pixel ones: {timeit(full_uniform):.1f} us")
    print(f"stage {stage}: conv + logits store {timeit(conv_only):.1f} us, + softmax/regression {timeit(no_fit):.1f} us, whole launch {timeit(full):.1f} us")
