import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from test_gpu_backward import case, make_module, cu
from mdf_net_b200 import ops
def timeit(fn, n=7, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(n):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts)//2]
stage=0
feats, ref_proj, src_projs, hyp, p, G, gout = case(stage, 576, 768, 5, 8, seed=12)
m = make_module(G, p).train()
rp = cu(ref_proj); sps = [cu(s) for s in src_projs]; hy = cu(hyp); go = cu(gout)
fs = [cu(f).requires_grad_(True) for f in feats]
fd = [f.detach() for f in fs]
cbr, fc = m.depth_weight[0], m.depth_weight[1]; bn = cbr.bn
def raw_fwd():
    return ops.cost_volume_train(fd, rp, sps, hy, cbr.conv.weight.detach(), bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, bn.eps, fc.weight.detach(), fc.bias.detach(), G, True)
out, stats = raw_fwd()
def raw_bwd():
    return ops.cost_volume_bwd(fd, rp, sps, hy, cbr.conv.weight.detach(), bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, bn.eps, fc.weight.detach(), fc.bias.detach(), G, True, out, go, stats)
print("raw train fwd op", timeit(raw_fwd), "ms; raw bwd op", timeit(raw_bwd), "ms")
def fwd(): return m(fs, rp, sps, hy)
def both(): m(fs, rp, sps, hy).backward(go)
print("module fwd", timeit(fwd), "ms; module fwd+bwd", timeit(both), "ms")
# graph-captured raw ops: pure GPU time
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    raw_fwd(); raw_bwd()
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    o2, s2 = raw_fwd(); r = raw_bwd()
print("graph replay raw fwd+bwd", timeit(lambda: g.replay()), "ms")
