#!/usr/bin/env python
"""Where does the end-to-end deviation of the drop-in model come from?  Reference CoreNet (oracle/_ref) with seeded weights:
teacher-forced per-stage comparison of VectorAggregate, then end-to-end with one unit swapped at a time."""
import contextlib
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import mdf_net_b200 as mdf
from mdf_net_b200 import synthetic as syn
from oracle import ref_install

torch.backends.cudnn.benchmark = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
with contextlib.redirect_stdout(io.StringIO()):
    r = ref_install.modules()
h0, w0, N = 512, 640, 3
ndepths, ngroups, curves, thresh = (48, 24, 8), (32, 16, 8), [None, "gauss1", "laplace"], (0.0, 0.95, 1e-5)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()


def build(agg_cls, depth_fn, conf_fn):
    torch.manual_seed(7)
    with contextlib.redirect_stdout(io.StringIO()):
        backbone = r.backbone.FPN_4Scales((8, 16, 32, 64))
        hypos = torch.nn.ModuleList([r.depthhypos.HyposByFit(ndepths[i], curves[i], thresh[i]) for i in range(3)])
        agg = torch.nn.ModuleList([agg_cls(g) for g in ngroups])
        reg = torch.nn.ModuleList([r.regular.RegularNet_3Scales(ngroups[0])] + [r.regular.RegularNet_4Scales(g) for g in ngroups[1:]])
        return r.core.CoreNet(backbone, hypos, r.scale.scale_cam, agg, reg, [depth_fn, conf_fn], r.refine.RefineNet2())


theirs = build(r.homoaggregate.VectorAggregate, r.regress.depth_regression, r.regress.confidence_regress)
K, E = syn.camera_rig(1, N, h0, w0, seed=5)
rng = np.random.default_rng(5)
imgs0 = cu(rng.random((1, N, 3, h0, w0), dtype=np.float32))
from oracle import ref_bench
theirs = ref_bench.randomise_weights(theirs.cuda(), imgs0[:, 0]).cpu()
sd = theirs.state_dict()
theirs = theirs.cuda().eval()
args = (imgs0, cu(E), cu(K), cu(np.array([[425.0, 935.0]], np.float32)))

captured = []
hooks = [m.register_forward_hook(lambda mod, inp, out: captured.append((inp, out))) for m in theirs.Homoaggre]
with torch.no_grad():
    a = theirs(*args)
for h in hooks:
    h.remove()
for s, (inp, out) in enumerate(captured):
    feats, ref_proj, src_projs, hyp = inp
    mine = mdf.VectorAggregate(ngroups[s]).cuda().eval()
    mine.load_state_dict(theirs.Homoaggre[s].state_dict())
    with torch.no_grad():
        got = mine(list(feats), ref_proj, list(src_projs), hyp)
    err = (got - out).abs()
    rel = float(torch.linalg.vector_norm((got - out).double()) / torch.linalg.vector_norm(out.double()))
    print(f"stage {s}: teacher-forced rel-L2 {rel:.3g}, max abs {float(err.max()):.3g}, >1e-5: {float((err > 1e-5).float().mean()):.2e}, "
          f">1e-4: {float((err > 1e-4).float().mean()):.2e}, >1e-3: {float((err > 1e-3).float().mean()):.2e}; features |max| {max(float(f.abs().max()) for f in feats):.3g}, "
          f"out range [{float(out.min()):.3f}, {float(out.max()):.3f}], hypos {tuple(hyp.shape)}")
    pix = err.amax(dim=(1, 2))[0]
    ys, xs = torch.nonzero(pix > 1e-4, as_tuple=True)
    if len(ys):
        print(f"         pixels with an element error > 1e-4: {len(ys)} (rows {int(ys.min())}..{int(ys.max())}, cols {int(xs.min())}..{int(xs.max())})")

within = lambda x, y: float(((x - y).abs() < 0.5).float().mean())
for name, (agg, dfn, cfn) in {"VectorAggregate only": (mdf.VectorAggregate, r.regress.depth_regression, r.regress.confidence_regress),
                              "regress only": (r.homoaggregate.VectorAggregate, mdf.depth_regression, mdf.confidence_regress),
                              "both": (mdf.VectorAggregate, mdf.depth_regression, mdf.confidence_regress)}.items():
    m = build(agg, dfn, cfn)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    with torch.no_grad():
        b = m(*args)
    e = (a["depth"] - b["depth"]).abs()
    print(f"{name}: depth within 0.5 mm {within(a['depth'], b['depth']):.5f}, median {float(e.median()):.3g} mm, p99 {float(e.flatten().kthvalue(int(0.99 * e.numel())).values):.3g}, max {float(e.max()):.3g}")
with torch.no_grad():
    a2 = theirs(*args)
print(f"reference vs itself: within 0.5 mm {within(a['depth'], a2['depth']):.5f}, max {float((a['depth'] - a2['depth']).abs().max()):.3g}")
for lvl in (1e-7, 1e-6, 1e-5):
    gen = torch.Generator(device="cuda").manual_seed(3)
    hooks = [m.register_forward_hook(lambda mod, inp, out: out * (1.0 + lvl * torch.randn(out.shape, device=out.device, generator=gen))) for m in theirs.Homoaggre]
    with torch.no_grad():
        c = theirs(*args)
    for h in hooks:
        h.remove()
    print(f"reference under {lvl:g} noise: within 0.5 mm {within(a['depth'], c['depth']):.5f}")
