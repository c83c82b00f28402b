#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs the read-only checkout at /root/reference and
torch CPU):   python tests/golden/make_golden.py

The reference ships no tests, golden vectors or checkpoints, so parity is
pinned on outputs of the reference's own functions imported as-is:
  net.unit.base.homo_warping, net.unit.homoaggregate.{VectorAggregate,
  homo_aggregate_by_variance}, net.unit.regress.{depth_regression, confidence_regress},
  net.unit.scale.scale_cam, net.unit.regular.RegularNet_{3,4}Scales (input / output of their last layer `.prob`),
  net.unit.depthhypos.HyposByFit, and a whole
  config.model (CoreNet) forward with per-stage tensors captured by hooks.
Inputs come from mdf_net_b200.synthetic (seeded); each .npz stores inputs and outputs so the
GPU box (which has no /root/reference) can replay them.
"""
import os
import sys

sys.dont_write_bytecode = True
REF = os.environ.get("MDF_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import numpy as np
import torch
import torch.nn.functional as F

from mdf_net_b200 import synthetic as syn

torch.set_num_threads(1)   # fixed reduction order inside ATen for the goldens
torch.manual_seed(1)

from net.unit.base import homo_warping                      # noqa: E402
from net.unit.homoaggregate import VectorAggregate, homo_aggregate_by_variance  # noqa: E402
from net.unit import regress, scale as ref_scale            # noqa: E402


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def op_rig(B, N, H, W, seed):
    """Projection matrices at feature resolution (as scale_cam would hand them over)."""
    K, E = syn.camera_rig(B, N, H, W, seed=seed)
    ref_proj, src_projs = ref_scale.scale_cam(T(K), T(E), 3)   # stage 3 -> level 0: no rescale
    return ref_proj, list(src_projs)


def set_depth_weight(mod, p):
    """Load synthetic depth_weight parameters into a reference VectorAggregate."""
    with torch.no_grad():
        mod.depth_weight[0].conv.weight.copy_(T(p["cw"]).view(1, -1, 1, 1, 1))
        bn = mod.depth_weight[0].bn
        bn.weight.fill_(float(p["bn_weight"])); bn.bias.fill_(float(p["bn_bias"]))
        bn.running_mean.fill_(float(p["bn_mean"])); bn.running_var.fill_(float(p["bn_var"]))
        mod.depth_weight[1].weight.fill_(float(p["fc_weight"])); mod.depth_weight[1].bias.fill_(float(p["fc_bias"]))


# ---------------------------------------------------------------------------- homo_warping
def gen_warp():
    for name, B, C, D, H, W, per_pixel, seed in [
        ("warp_uniform", 2, 8, 6, 12, 16, False, 11),
        ("warp_pixel", 2, 6, 5, 10, 14, True, 12),
    ]:
        ref_proj, src_projs = op_rig(B, 3, H, W, seed)
        fea = syn.smooth_features(B, 1, C, H, W, seed=seed)[0]
        hyp = syn.pixel_hypos(B, D, H, W, seed=seed) if per_pixel else syn.uniform_hypos(B, D)
        outs, projs = [], []
        for sp in src_projs:
            outs.append(homo_warping(T(fea), sp, ref_proj, T(hyp)).numpy())
            projs.append(torch.matmul(sp, torch.inverse(ref_proj)).numpy())   # base.py:98
        save(name, src_fea=fea, ref_proj=ref_proj.numpy(), src_projs=np.stack([s.numpy() for s in src_projs]),
             depth_hypos=hyp, proj=np.stack(projs), warped=np.stack(outs))

    # edge cases: hypotheses that put points behind / on the source camera plane, huge and
    # non-finite depths, and the identity homography (documents the half-pixel shift)
    B, C, D, H, W = 1, 4, 8, 9, 11
    ref_proj, src_projs = op_rig(B, 2, H, W, 13)
    fea = syn.smooth_features(B, 1, C, H, W, seed=13)[0]
    hyp = np.array([425.0, -300.0, 0.0, 1e-3, 1e9, np.inf, np.nan, 935.0], np.float32).reshape(1, D, 1, 1)
    out = homo_warping(T(fea), src_projs[0], ref_proj, T(hyp)).numpy()
    ident = homo_warping(T(fea), ref_proj, ref_proj, T(hyp[:, :1])).numpy()
    save("warp_edge", src_fea=fea, ref_proj=ref_proj.numpy(), src_proj=src_projs[0].numpy(), depth_hypos=hyp,
         warped=out, warped_identity=ident)


# ------------------------------------------------------------------------- VectorAggregate
def gen_vecagg():
    for name, B, N, C, G, D, H, W, per_pixel, seed in [
        ("vecagg_s0", 1, 3, 64, 32, 48, 6, 8, False, 21),
        ("vecagg_s1", 2, 5, 32, 16, 24, 8, 8, True, 22),
        ("vecagg_s2", 1, 4, 16, 8, 8, 16, 24, True, 23),
        ("vecagg_cpg4", 1, 3, 16, 4, 6, 8, 10, False, 24),
        ("vecagg_n2", 2, 2, 16, 8, 4, 7, 9, True, 25),     # ragged sizes, single source view
    ]:
        ref_proj, src_projs = op_rig(B, N, H, W, seed)
        feats = syn.smooth_features(B, N, C, H, W, seed=seed)
        hyp = syn.pixel_hypos(B, D, H, W, seed=seed) if per_pixel else syn.uniform_hypos(B, D)
        p = syn.depth_weight_params(G, seed=seed)
        mod = VectorAggregate(G).eval()
        set_depth_weight(mod, p)
        with torch.no_grad():
            out = mod([T(f) for f in feats], ref_proj, src_projs, T(hyp)).numpy()
        save(name, features=np.stack(feats), ref_proj=ref_proj.numpy(),
             src_projs=np.stack([s.numpy() for s in src_projs]), depth_hypos=hyp, groups=G,
             cost_volume=out, **{"p_" + k: v for k, v in p.items()})


def gen_varagg():
    B, N, C, D, H, W, seed = 1, 3, 16, 8, 8, 12, 31
    ref_proj, src_projs = op_rig(B, N, H, W, seed)
    feats = syn.smooth_features(B, N, C, H, W, seed=seed)
    for name, hyp in [("varagg_uniform", syn.uniform_hypos(B, D)), ("varagg_pixel", syn.pixel_hypos(B, D, H, W, seed=seed))]:
        out = homo_aggregate_by_variance([T(f) for f in feats], ref_proj, src_projs, T(hyp)).numpy()
        save(name, features=np.stack(feats), ref_proj=ref_proj.numpy(),
             src_projs=np.stack([s.numpy() for s in src_projs]), depth_hypos=hyp, cost_volume=out)


def gen_vecagg_grad():
    """Gradients of the unmodified reference VectorAggregate (autograd) in eval and train mode: w.r.t. every
    feature map and the depth_weight parameters, for a fixed upstream gradient; train mode also records the
    running statistics after the forward (BatchNorm3d updates them once per source view)."""
    for name, B, N, C, G, D, H, W, per_pixel, seed in [
        ("vecagg_grad_s0", 2, 3, 64, 32, 6, 6, 8, False, 71),
        ("vecagg_grad_s2", 1, 4, 16, 8, 8, 10, 12, True, 72),
    ]:
        ref_proj, src_projs = op_rig(B, N, H, W, seed)
        feats = syn.smooth_features(B, N, C, H, W, seed=seed)
        hyp = syn.pixel_hypos(B, D, H, W, seed=seed) if per_pixel else syn.uniform_hypos(B, D)
        p = syn.depth_weight_params(G, seed=seed)
        gout = np.random.default_rng(seed).standard_normal((B, G, D, H, W)).astype(np.float32)
        out = {}
        for mode in ("eval", "train"):
            mod = VectorAggregate(G)
            set_depth_weight(mod, p)
            mod.train(mode == "train")
            fs = [T(f).clone().requires_grad_(True) for f in feats]
            cv = mod(fs, ref_proj, src_projs, T(hyp))
            cv.backward(T(gout))
            out[f"{mode}_cost_volume"] = cv.detach().numpy()
            out[f"{mode}_grad_features"] = np.stack([f.grad.numpy() for f in fs])
            dw = mod.depth_weight
            out[f"{mode}_grad_cw"] = dw[0].conv.weight.grad.numpy().reshape(-1)
            out[f"{mode}_grad_bn"] = np.array([dw[0].bn.weight.grad.item(), dw[0].bn.bias.grad.item()], np.float32)
            out[f"{mode}_grad_fc"] = np.array([dw[1].weight.grad.item(), dw[1].bias.grad.item()], np.float32)
            out[f"{mode}_running"] = np.array([dw[0].bn.running_mean.item(), dw[0].bn.running_var.item(),
                                               float(dw[0].bn.num_batches_tracked.item())], np.float64)
        save(name, features=np.stack(feats), ref_proj=ref_proj.numpy(), src_projs=np.stack([s.numpy() for s in src_projs]),
             depth_hypos=hyp, groups=G, grad_out=gout, **{"p_" + k: v for k, v in p.items()}, **out)


# ------------------------------------------------------------------------------------ head
def gen_head():
    for name, B, D, H, W, seed in [("head_d48", 1, 48, 8, 12, 41), ("head_d24", 2, 24, 8, 12, 42), ("head_d8", 2, 8, 12, 16, 43)]:
        logits = syn.regulariser_logits(B, D, H, W, seed=seed)
        prob = F.softmax(T(logits), dim=1)                                   # regular.py:69,133
        hyp_u, hyp_p = syn.uniform_hypos(B, D), syn.pixel_hypos(B, D, H, W, seed=seed)
        depth_u = regress.depth_regression(prob, T(hyp_u)).numpy()
        depth_p = regress.depth_regression(prob, T(hyp_p)).numpy()
        conf = regress.confidence_regress(prob)
        conf_up = F.interpolate(conf.unsqueeze(1), size=None, scale_factor=2, mode="nearest",
                                align_corners=None).squeeze(1)              # core.py:76-77
        last = T(np.random.default_rng(seed).uniform(0, 1, (B, H // 2, W // 2)).astype(np.float32))
        conf_blend = regress.confidence_regress(prob, last_confidence=last)
        save(name, logits=logits, prob=prob.numpy(), hypos_uniform=hyp_u, hypos_pixel=hyp_p,
             depth_uniform=depth_u, depth_pixel=depth_p, confidence=conf.numpy(), confidence_up=conf_up.numpy(),
             last_confidence=last.numpy(), confidence_blend=conf_blend.numpy())
    # known answers: one-hot and uniform probability volumes
    D, H, W = 8, 2, 8
    onehot = np.zeros((1, D, H, W), np.float32)
    for i in range(H * W):
        onehot[0, i % D, i // W, i % W] = 1.0
    unif = np.full((1, D, H, W), 1.0 / D, np.float32)
    save("head_known", onehot=onehot, onehot_conf=regress.confidence_regress(T(onehot)).numpy(),
         uniform=unif, uniform_conf=regress.confidence_regress(T(unif)).numpy())


# ------------------------------------------------------------------------------ HyposByFit
def gen_hypos():
    """HyposByFit (net/unit/depthhypos.py) as config.py:199-203 wires it: stage 0 uniform, stage 0 -> 1 'gauss1'
    (prob_thresh 0.95), stage 1 -> 2 'laplace' (prob_thresh 1e-5), both with the x2 upsampling core.py:55 asks for.
    The float64 run of the same reference code is stored next to the float32 one: the gauss1 normal equations
    are so ill conditioned in float32 that the reference's own result is only good to ~1e-3."""
    from net.unit.depthhypos import HyposByFit
    B, H, W = 2, 12, 16
    dr = np.array([[425.0, 935.0], [480.0, 900.0]], np.float32)
    out = {"depth_range": dr}
    m0 = HyposByFit(48, None, 0.0)
    hyp0 = m0(None, T(dr), None, None, upsample=True)
    out["hypos0"] = hyp0.numpy()
    prob0 = F.softmax(T(syn.regulariser_logits(B, 48, H, W, seed=81, peak=6.0)), 1)
    depth0 = regress.depth_regression(prob0, hyp0)
    m1 = HyposByFit(24, "gauss1", 0.95)
    hyp1 = m1(depth0, T(dr), prob0, hyp0, upsample=True)
    hyp1_64 = m1(depth0.double(), T(dr).double(), prob0.double(), hyp0.double(), upsample=True)
    out.update(prob0=prob0.numpy(), depth0=depth0.numpy(), hypos1=hyp1.numpy(), hypos1_f64=hyp1_64.numpy(),
               s1=m1._gauss_fitting1(depth0, prob0, hyp0).numpy(),
               s1_f64=m1._gauss_fitting1(depth0.double(), prob0.double(), hyp0.double()).numpy())
    prob1 = F.softmax(T(syn.regulariser_logits(B, 24, 2 * H, 2 * W, seed=82, peak=6.0)), 1)
    depth1 = regress.depth_regression(prob1, hyp1)
    m2 = HyposByFit(8, "laplace", 1e-5)
    hyp2 = m2(depth1, T(dr), prob1, hyp1, upsample=True)
    hyp2_64 = m2(depth1.double(), T(dr).double(), prob1.double(), hyp1.double(), upsample=True)
    out.update(prob1=prob1.numpy(), depth1=depth1.numpy(), hypos2=hyp2.numpy(), hypos2_f64=hyp2_64.numpy(),
               s2=m2._laplace_fitting(depth1, prob1, hyp1).numpy(),
               s2_f64=m2._laplace_fitting(depth1.double(), prob1.double(), hyp1.double()).numpy())
    save("hypos_fit", **out)


def gen_hypos_gauss0():
    """The curve config.py does not wire: HyposByFit(curve 'gauss0') (depthhypos.py:42-43, 127-167) on uniform (stage 0 -> 1)
    and on per-pixel hypotheses, float32 and float64 runs of the reference (its 2x2 normal equations are as ill conditioned
    in float32 as gauss1's)."""
    from net.unit.depthhypos import HyposByFit
    B, H, W = 2, 12, 16
    dr = np.array([[425.0, 935.0], [480.0, 900.0]], np.float32)
    out = {"depth_range": dr}
    hyp0 = HyposByFit(48, None, 0.0)(None, T(dr), None, None, upsample=True)
    prob0 = F.softmax(T(syn.regulariser_logits(B, 48, H, W, seed=83, peak=6.0)), 1)
    depth0 = regress.depth_regression(prob0, hyp0)
    m = HyposByFit(24, "gauss0", 0.95)
    out.update(hypos0=hyp0.numpy(), prob0=prob0.numpy(), depth0=depth0.numpy(),
               s=m._gauss_fitting0(depth0, prob0, hyp0).numpy(),
               s_f64=m._gauss_fitting0(depth0.double(), prob0.double(), hyp0.double()).numpy(),
               hypos1=m(depth0, T(dr), prob0, hyp0, upsample=True).numpy(),
               hypos1_f64=m(depth0.double(), T(dr).double(), prob0.double(), hyp0.double(), upsample=True).numpy())
    hyp_p = T(syn.pixel_hypos(B, 24, H, W, seed=84))
    prob_p = F.softmax(T(syn.regulariser_logits(B, 24, H, W, seed=85, peak=6.0)), 1)
    depth_p = regress.depth_regression(prob_p, hyp_p)
    m8 = HyposByFit(8, "gauss0", 0.9)
    out.update(hypos_p=hyp_p.numpy(), prob_p=prob_p.numpy(), depth_p=depth_p.numpy(),
               s_p=m8._gauss_fitting0(depth_p, prob_p, hyp_p).numpy(),
               s_p_f64=m8._gauss_fitting0(depth_p.double(), prob_p.double(), hyp_p.double()).numpy(),
               hypos2_f64=m8(depth_p.double(), T(dr).double(), prob_p.double(), hyp_p.double(), upsample=False).numpy())
    save("hypos_fit_gauss0", **out)


# ------------------------------------------------------ regulariser tail: prob conv + softmax + head + fit
def gen_prob_head():
    """The last layer of the reference's regularisers (net/unit/regular.py:43,67-69 and :110,130-133) with what follows
    it in CoreNet.forward: `self.prob` (Conv3d(c0,1,3,pad=1,no bias)) -> softmax -> depth_regression (-> confidence on the
    last stage) (-> the curve fit of the next stage's HyposByFit).  The unmodified modules run on a random cost volume;
    hooks capture the input and output of `.prob`."""
    from net.unit.depthhypos import HyposByFit
    from net.unit.regular import RegularNet_3Scales, RegularNet_4Scales
    g = torch.Generator().manual_seed(97)
    for name, net, B, G, D, H, W, curve, seed in [
        ("prob_head_s0", RegularNet_3Scales(32), 1, 32, 48, 8, 12, "gauss1", 101),
        ("prob_head_s1", RegularNet_4Scales(16), 2, 16, 24, 16, 24, "laplace", 102),
        ("prob_head_s2", RegularNet_4Scales(8), 1, 8, 8, 16, 24, None, 103),
    ]:
        net = net.eval()
        with torch.no_grad():
            for m in net.modules():
                if isinstance(m, torch.nn.BatchNorm3d):
                    m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                    m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
                    m.weight.copy_(0.8 + 0.4 * torch.rand(m.weight.shape, generator=g))
                    m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
        cv = T(np.random.default_rng(seed).uniform(0.2, 0.8, (B, G, D, H, W)).astype(np.float32))
        cap = {}
        h1 = net.prob.register_forward_pre_hook(lambda mod, args: cap.__setitem__("x", args[0].numpy().copy()))
        h2 = net.prob.register_forward_hook(lambda mod, args, out: cap.__setitem__("logits", out.squeeze(1).numpy().copy()))
        with torch.no_grad():
            net(cv)
            net.prob.weight.mul_(2.0 / float(np.std(cap["logits"])))       # useful logits (default init gives ~1e-2)
            prob = net(cv)
        h1.remove(); h2.remove()
        hyp = syn.uniform_hypos(B, D) if name.endswith("s0") else syn.pixel_hypos(B, D, H, W, seed=seed)
        depth = regress.depth_regression(prob, T(hyp))
        out = dict(x=cap["x"], weight=net.prob.weight.detach().numpy(), logits=cap["logits"], prob=prob.numpy(),
                   depth_hypos=hyp, depth=depth.numpy())
        if curve is None:
            conf = regress.confidence_regress(prob)
            out["confidence_up"] = F.interpolate(conf.unsqueeze(1), size=None, scale_factor=2, mode="nearest",
                                                 align_corners=None).squeeze(1).numpy()       # core.py:76-77
        else:
            m = HyposByFit(8, curve, 0.95)
            fit = m._gauss_fitting1 if curve == "gauss1" else m._laplace_fitting
            out["s"] = fit(depth, prob, T(hyp)).numpy()
            out["s_f64"] = fit(depth.double(), prob.double(), T(hyp).double()).numpy()
        save(name, **out)


# ------------------------------------------------------------------ geometric-consistency filter (post-processing)
def filter_scene(S, H, W, seed):
    """A tilted plane seen by S+1 cameras of the synthetic rig, with smooth multiplicative noise and a few outliers in
    every depth map, so that every dynamic threshold of the filter gets both outcomes."""
    rng = np.random.default_rng(seed)
    K, E = syn.camera_rig(1, S + 1, H, W, seed=seed)
    K, E = K[0].astype(np.float64), E[0].astype(np.float64)
    R0, t0 = E[0, :3, :3], E[0, :3, 3]
    p0 = R0.T @ (np.array([0.0, 0.0, 700.0]) - t0)               # 700 mm in front of the reference camera
    n = R0.T @ np.array([0.15, -0.1, 1.0]); n /= np.linalg.norm(n)
    c = float(n @ p0)
    ys, xs = np.mgrid[0:H, 0:W].astype(np.float64)
    depths = []
    for v in range(S + 1):
        R, t = E[v, :3, :3], E[v, :3, 3]
        dirs = np.einsum("ij,jhw->ihw", np.linalg.inv(K[v]), np.stack([xs, ys, np.ones_like(xs)]))
        num = c + n @ (R.T @ t)
        den = np.einsum("i,ihw->hw", n @ R.T, dirs)
        d = num / den
        noise = syn._upsample2x_bilinear(syn._upsample2x_bilinear(rng.standard_normal((1, 1, H // 4, W // 4)).astype(np.float32)))[0, 0]
        d = d * (1.0 + 0.004 * noise[:H, :W])
        out = rng.random((H, W)) < 0.04
        d = np.where(out, d * rng.uniform(0.8, 1.2, (H, W)), d)
        depths.append(d.astype(np.float32))
    conf = rng.uniform(0.5, 1.0, (H, W)).astype(np.float32)
    return K.astype(np.float32), E.astype(np.float32), depths, conf


def gen_geo_filter():
    """tools/filter/dynamic_filter_gpu.py: check_geometric_consistency / reproject_with_depth imported as they are (the
    module's `plyfile` import, only used by its PLY writer, is stubbed) and run on CPU tensors; the per-view aggregation
    is lines 78-98 of filter() (which itself needs CUDA, the dataset on disk and plyfile)."""
    import types
    sys.modules.setdefault("plyfile", types.SimpleNamespace(PlyData=None, PlyElement=None))
    sys.path.insert(0, os.path.join(REF, "tools", "filter"))
    saved = {k: os.environ.get(k) for k in ("CUDA_VISIBLE_DEVICES", "CUDA_DEVICE_ORDER")}
    import dynamic_filter_gpu as dfg
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    S, H, W = 4, 96, 128
    K, E, depths, conf = filter_scene(S, H, W, seed=111)
    thre1, thre2, photo_threshold, nconditions = 4, 1300.0, 0.8, 2
    ref_depth = T(depths[0])
    avg_mask, reproj, sums, bits = 0, [], None, []
    for v in range(1, S + 1):
        masks, mask, drep = dfg.check_geometric_consistency(ref_depth, T(K[0]), T(E[0]), T(depths[v]), T(K[v]), T(E[v]), thre1, thre2)
        masks = [m.float() for m in masks]
        sums = masks if sums is None else [a + b for a, b in zip(sums, masks)]
        avg_mask = avg_mask + mask
        reproj.append(drep)
        bits.append(sum((m[0].numpy().astype(np.uint16) << i) for i, m in enumerate(masks)))
    geo = 0
    for i in range(2, 11):
        geo = geo + (sums[i - 2] >= i)
    averaged = (sum(reproj) + ref_depth) / (avg_mask + 1)
    geo = geo >= nconditions
    photo = T(conf) > photo_threshold
    final = torch.logical_and(photo, geo)
    save("geo_filter", intrinsics=K, extrinsics=E, depths=np.stack(depths), confidence=conf,
         params=np.array([thre1, thre2, photo_threshold, nconditions], np.float64),
         bits=np.stack(bits), depth_reprojected=np.stack([r[0].numpy() for r in reproj]),
         depth_averaged=averaged[0].numpy(), geo=geo[0].numpy(), photo=photo.numpy(), final=final[0].numpy())


# -------------------------------------------------------------------------------- scale_cam
def gen_scale():
    K, E = syn.camera_rig(2, 4, 64, 80, seed=51)
    out = {}
    for stage in range(3):
        ref_proj, src_projs = ref_scale.scale_cam(T(K), T(E), stage)
        out[f"ref_proj_{stage}"] = ref_proj.numpy()
        out[f"src_projs_{stage}"] = np.stack([s.numpy() for s in src_projs])
    save("scale_cam", intrinsics=K, extrinsics=E, **out)


# ---------------------------------------------------------------------- whole CoreNet forward
def gen_corenet():
    import config  # builds config.model at import (config.py:186-218)
    model = config.model.eval()
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
                m.weight.copy_(0.8 + 0.4 * torch.rand(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
        for i, hm in enumerate(model.Homoaggre):
            set_depth_weight(hm, syn.depth_weight_params(hm.ngroups, seed=60 + i))

    B, N, H0, W0 = 1, 3, 64, 64
    rng = np.random.default_rng(61)
    imgs = rng.uniform(0, 1, (B, N, 3, H0, W0)).astype(np.float32)
    K, E = syn.camera_rig(B, N, H0, W0, seed=61)
    depth_range = np.array([[425.0, 935.0]], np.float32)

    with torch.no_grad():
        # default-init features are ~1e-4 and give a degenerate cost volume: rescale the
        # FPN output convs to unit-ish feature magnitude, and the prob convs to useful logits
        feats = model.Backbone(T(imgs[:, 0]))
        for conv, f in zip((model.Backbone.out4, model.Backbone.out3, model.Backbone.out2), feats):
            conv.weight.mul_(1.5 / float(f.std()))

    cap = {}

    def hook_homo(i):
        def fn(mod, args, out):
            features, ref_proj, src_projs, hyp = args
            cap[f"s{i}_features"] = np.stack([f.numpy() for f in features])
            cap[f"s{i}_ref_proj"] = ref_proj.numpy()
            cap[f"s{i}_src_projs"] = np.stack([s.numpy() for s in src_projs])
            cap[f"s{i}_depth_hypos"] = hyp.numpy()
            cap[f"s{i}_cost_volume"] = out.numpy()
            p = mod.depth_weight
            cap[f"s{i}_cw"] = p[0].conv.weight.detach().numpy().reshape(-1)
            cap[f"s{i}_bn"] = np.array([p[0].bn.weight.item(), p[0].bn.bias.item(), p[0].bn.running_mean.item(),
                                        p[0].bn.running_var.item(), p[0].bn.eps], np.float64)
            cap[f"s{i}_fc"] = np.array([p[1].weight.item(), p[1].bias.item()], np.float64)
        return fn

    def hook_prob_conv(i):
        def fn(mod, args, out):
            cap[f"s{i}_logits"] = out.squeeze(1).numpy()
        return fn

    def hook_regular(i):
        def fn(mod, args, out):
            cap[f"s{i}_prob"] = out.numpy()
        return fn

    with torch.no_grad():
        # first pass to find the logit scale, then rescale prob convs
        handles = [model.Regular[i].prob.register_forward_hook(hook_prob_conv(i)) for i in range(3)]
        model(T(imgs), T(E), T(K), T(depth_range))
        for i in range(3):
            model.Regular[i].prob.weight.mul_(2.0 / float(np.std(cap[f"s{i}_logits"])))
        handles += [model.Homoaggre[i].register_forward_hook(hook_homo(i)) for i in range(3)]
        handles += [model.Regular[i].register_forward_hook(hook_regular(i)) for i in range(3)]
        depths = []
        orig = model.Depth_regress
        model.Depth_regress = lambda p, h: depths.append(orig(p, h)) or depths[-1]
        out = model(T(imgs), T(E), T(K), T(depth_range))
        model.Depth_regress = orig
        for h in handles:
            h.remove()
    for i, d in enumerate(depths):
        cap[f"s{i}_depth"] = d.numpy()
    cap["confidence"] = out["confidence"].numpy()
    cap["final_depth"] = out["depth"].numpy()
    cap["intrinsics"], cap["extrinsics"], cap["depth_range"] = K, E, depth_range
    keys = sorted(model.Homoaggre.state_dict().keys())
    cap["homoaggre_state_keys"] = np.array(keys)
    save("corenet_64x64_n3", **cap)


def gen_output_files():
    """The files eval.py:36-50 writes per view, byte for byte: depth PFM, depth PNG ((d-500)/2 as 8-bit), confidence PFM
    (tools/data_io.py:44-75), on a small seeded depth / confidence pair."""
    import tempfile
    from tools.data_io import save_pfm, write_depth_img
    rng = np.random.default_rng(12)
    depth = (700.0 + 150.0 * rng.standard_normal((48, 64))).astype(np.float32)      # some values leave [500, 1010]: the PNG clips
    conf = rng.random((48, 64), dtype=np.float32)
    with tempfile.TemporaryDirectory() as d:
        save_pfm(os.path.join(d, "d.pfm"), torch.from_numpy(depth))
        write_depth_img(os.path.join(d, "d.png"), depth)
        save_pfm(os.path.join(d, "c.pfm"), torch.from_numpy(conf))
        rd = lambda n: np.frombuffer(open(os.path.join(d, n), "rb").read(), np.uint8)
        save("output_files", depth=depth, confidence=conf, depth_pfm=rd("d.pfm"), depth_png=rd("d.png"), confidence_pfm=rd("c.pfm"))


if __name__ == "__main__":
    only = sys.argv[1:]
    for fn in (gen_output_files, gen_warp, gen_vecagg, gen_vecagg_grad, gen_varagg, gen_head, gen_hypos, gen_hypos_gauss0, gen_prob_head, gen_geo_filter, gen_scale, gen_corenet):
        if not only or fn.__name__ in only:
            fn()
