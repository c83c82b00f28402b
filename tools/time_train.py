import os, sys
ROOT = "/root/repo" if os.path.isdir("/root/repo/tests") else os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from test_gpu_backward import case, make_module, cu
def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(n):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts)//2]
for stage in range(3):
    feats, ref_proj, src_projs, hyp, p, G, gout = case(stage, 576, 768, 5, 8, seed=12)
    m = make_module(G, p).train()
    rp = cu(ref_proj); sps = [cu(s) for s in src_projs]; hy = cu(hyp); go = cu(gout)
    fs = [cu(f).requires_grad_(True) for f in feats]
    def fwd(): return m(fs, rp, sps, hy)
    def both(): m(fs, rp, sps, hy).backward(go)
    with torch.no_grad():
        me = make_module(G, p).eval()
        t_eval = timeit(lambda: me([f.detach() for f in fs], rp, sps, hy))
    t_f, t_b = timeit(fwd), timeit(both)
    t_s = timeit(lambda: [both() for _ in range(4)]) / 4.0
    print(f"stage {stage}: eval fwd {t_eval:.2f} ms, train fwd {t_f:.2f} ms, fwd+bwd {t_b:.2f} ms (steady state, 4 iterations back to back: {t_s:.2f} ms)")
