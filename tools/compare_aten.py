#!/usr/bin/env python
"""The same GPU, the same inputs: this repo's fused ops vs a plain-PyTorch (ATen eager) evaluation of the reference's
formulae (tests/torch_ref.py, pinned on the reference's outputs).  Forward at BASELINE configs[1] shapes, and
forward+backward at the BlendedMVS train shape (configs[4]).  Run on a GPU box: python tools/compare_aten.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import torch_ref
import mdf_net_b200 as mdf
from test_gpu_backward import case, make_module, cu


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


print("== eval forward, 1600x1152 N=5 B=1 (ms per stage): fused op vs ATen eager ==")
tot_f = tot_a = 0.0
for stage in range(3):
    feats, ref_proj, src_projs, hyp, p, G, _ = case(stage, 1152, 1600, 5, 1, seed=11)
    m = make_module(G, p).eval()
    fs = [cu(f) for f in feats]; rp = cu(ref_proj); sps = [cu(s) for s in src_projs]; hy = cu(hyp)
    P = {k: cu(np.asarray(p[k], np.float32).reshape(-1)) for k in ("cw", "bn_weight", "bn_bias", "fc_weight", "fc_bias")}
    with torch.no_grad():
        t_f = timeit(lambda: m(fs, rp, sps, hy))
        t_a = timeit(lambda: torch_ref.vector_aggregate(fs, rp, sps, hy, P["cw"], P["bn_weight"], P["bn_bias"], float(p["bn_mean"]),
                                                        float(p["bn_var"]), float(p["bn_eps"]), P["fc_weight"], P["fc_bias"], G)[0])
    tot_f += t_f; tot_a += t_a
    print(f"stage {stage}: fused {t_f:.3f}  ATen eager {t_a:.3f}  x{t_a / t_f:.1f}")
print(f"sum: fused {tot_f:.3f}  ATen eager {tot_a:.3f}  x{tot_a / tot_f:.1f}")

print("== train forward + backward, 768x576 N=5 B=8 (ms per stage): CUDA autograd path vs ATen autograd ==")
for stage in range(3):
    feats, ref_proj, src_projs, hyp, p, G, gout = case(stage, 576, 768, 5, 8, seed=12)
    m = make_module(G, p).train()
    rp = cu(ref_proj); sps = [cu(s) for s in src_projs]; hy = cu(hyp); go = cu(gout)
    fs = [cu(f).requires_grad_(True) for f in feats]
    P = {k: cu(np.asarray(p[k], np.float32).reshape(-1)).requires_grad_(True) for k in ("cw", "bn_weight", "bn_bias", "fc_weight", "fc_bias")}

    def ours():
        m(fs, rp, sps, hy).backward(go)

    def aten():
        torch_ref.vector_aggregate(fs, rp, sps, hy, P["cw"], P["bn_weight"], P["bn_bias"], float(p["bn_mean"]), float(p["bn_var"]),
                                   float(p["bn_eps"]), P["fc_weight"], P["fc_bias"], G, training=True)[0].backward(go)
    t_o, t_a = timeit(ours, 3, 1), timeit(aten, 3, 1)
    print(f"stage {stage}: cuda {t_o:.2f}  ATen autograd {t_a:.2f}  x{t_a / t_o:.1f}   peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
