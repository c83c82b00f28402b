"""Drop-in replacements for the reference's hot-path units, same names and call signatures.

They plug into the reference's own injection point, `core.CoreNet(Backbone, Depth_hypos, scale,
Homoaggre, Regular, Regress, Refine)` (net/core.py:5-26, wired in config.py:186-218):

    Homoaggre = nn.ModuleList([mdf_net_b200.VectorAggregate(g) for g in (32, 16, 8)])
    Regress   = [mdf_net_b200.depth_regression, mdf_net_b200.confidence_regress]

`VectorAggregate` keeps the reference's state-dict keys (`depth_weight.0.conv.weight`,
`depth_weight.0.bn.{weight,bias,running_mean,running_var,num_batches_tracked}`,
`depth_weight.1.{weight,bias}`), so `load_state_dict(strict=True)` of a reference checkpoint works
(eval.py:15-17).  All compute runs in libmdf_b200.so; tensors must live on a CUDA device.
"""
from __future__ import annotations

import warnings
from typing import Sequence

import torch
import torch.nn as nn
from torch import Tensor

from . import ops


class _PointwiseConvBNReLU3D(nn.Module):
    """Parameter container with the sub-module names of the reference's ConvBNReLU3D
    (net/unit/base.py:50-68) for the k=1 case used by depth_weight; never called as a layer."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.conv = nn.Conv3d(in_channels, out_channels, 1, 1, 0, bias=False)
        self.bn = nn.BatchNorm3d(out_channels)
        self.relu = nn.ReLU(inplace=True)


class VectorAggregate(nn.Module):
    """Fused homo_warping + group-softmax similarity + learned view weighting.

    Mirrors net/unit/homoaggregate.py:8-46.  forward(features, ref_proj, src_projs, depth_hypos)
    -> cost volume (B, ngroups, D, H, W).
    """

    def __init__(self, ngroups: int = 8, algo: int = 0):
        super().__init__()
        self.ngroups = ngroups
        self.algo = algo
        # (B,G,D,H,W) -> (B,1,D,H,W): Conv3d(G,1,k=1,no bias) - BN - ReLU - Conv3d(1,1,k=1) - Sigmoid
        self.depth_weight = nn.Sequential(_PointwiseConvBNReLU3D(ngroups, 1), nn.Conv3d(1, 1, 1, 1, 0), nn.Sigmoid())

    def forward(self, features: Sequence[Tensor], ref_proj: Tensor, src_projs: Sequence[Tensor],
                depth_hypos: Tensor) -> Tensor:
        cbr, fc = self.depth_weight[0], self.depth_weight[1]
        if isinstance(features, PreppedFeatures):          # the FPN hand-off (inference only): no layout pass
            if self.training:
                raise RuntimeError("mdfnet_b200: prepared features are an inference path (no backward through the hand-off)")
            return ops.cost_volume_prepped(features.s4, features.q4, features.cq4, ref_proj, list(src_projs), depth_hypos,
                                           cbr.conv.weight, cbr.bn.weight, cbr.bn.bias, cbr.bn.running_mean, cbr.bn.running_var,
                                           cbr.bn.eps, fc.weight, fc.bias, self.ngroups)
        needs_grad = torch.is_grad_enabled() and any(t.requires_grad for t in (*features, *self.parameters()))
        if self.training or needs_grad:
            # train-mode BatchNorm (batch statistics per source view) and / or autograd: csrc/mdf_backward.cu.
            # (Inference should run under torch.no_grad(), as eval.py:24 does: an eval-mode call with grad enabled and
            # parameters that require grad takes this differentiable path, which saves the inputs for backward.)
            C = features[0].shape[1]
            if C != 2 * self.ngroups or self.ngroups not in (8, 16, 32):
                if self.training:
                    raise NotImplementedError(
                        f"mdfnet_b200: train-mode VectorAggregate needs C == 2*G with G in (8, 16, 32) (the reference's configured "
                        f"stages, config.py:196-205); got C={C}, G={self.ngroups}")
                warnings.warn(f"mdfnet_b200: no backward kernel for C={C}, G={self.ngroups}: this eval-mode call runs the "
                              "non-differentiable kernel (wrap inference in torch.no_grad())", RuntimeWarning, stacklevel=2)
            else:
                from . import autograd
                return autograd.vector_aggregate_train(self, list(features), ref_proj, list(src_projs), depth_hypos)
        return ops.cost_volume(list(features), ref_proj, list(src_projs), depth_hypos,
                               cbr.conv.weight, cbr.bn.weight, cbr.bn.bias, cbr.bn.running_mean, cbr.bn.running_var,
                               cbr.bn.eps, fc.weight, fc.bias, self.ngroups, self.algo)


class PreppedFeatures:
    """What `FPNHandOff` hands to `VectorAggregate` instead of the list of NCHW feature maps: the hot kernel's own input
    layout for one stage (q4 / cq4 of the reference view, s4 of the source views back to back)."""

    def __init__(self, q4: Tensor, cq4: Tensor, s4: Tensor):
        self.q4, self.cq4, self.s4 = q4, cq4, s4


class FPNHandOff(nn.Module):
    """Optional fast entry.  Wraps a backbone shaped like the reference's FPN_4Scales
    (net/unit/backbone.py:9-66: trunk conv01..conv34, laterals lat2 / lat3, bias-free 1x1 output convolutions out4 / out3 /
    out2) and runs it up to -- not including -- those output convolutions; `ops.fpn_out_prepped` then applies them and
    writes the pair-difference maps the cost-volume kernel gathers from, so neither the NCHW features nor the layout pass
    exist on this path.  The wrapped module is used as it is (same parameters, same state-dict keys).

    forward(views, aggregates) -> [PreppedFeatures per stage]; views: the N images, view 0 = reference (core.py:39-42);
    aggregates: the VectorAggregate of every stage (their depth_weight.0.conv.weight goes into cq4)."""

    def __init__(self, backbone: nn.Module):
        super().__init__()
        self.backbone = backbone

    @staticmethod
    def supports(backbone: nn.Module) -> bool:
        names = ("conv01", "conv12", "conv23", "conv34", "lat2", "lat3", "out2", "out3", "out4")
        if not all(hasattr(backbone, n) for n in names):
            return False
        return all(isinstance(c, nn.Conv2d) and c.bias is None and tuple(c.kernel_size) == (1, 1) and tuple(c.stride) == (1, 1)
                   and c.groups == 1 and c.in_channels % 4 == 0 for c in (backbone.out2, backbone.out3, backbone.out4))

    def trunk(self, x: Tensor):
        """backbone.py:51-63 without the three output convolutions: the merged maps at 1/8, 1/4, 1/2."""
        bb = self.backbone
        x2 = bb.conv12(bb.conv01(x))
        x3 = bb.conv23(x2)
        x4 = bb.conv34(x3)
        m3 = nn.functional.interpolate(x4, scale_factor=2.0, mode="bilinear", align_corners=False) + bb.lat3(x3)
        m2 = nn.functional.interpolate(m3, scale_factor=2.0, mode="bilinear", align_corners=False) + bb.lat2(x2)
        return x4, m3, m2

    def forward(self, views: Sequence[Tensor], aggregates: Sequence["VectorAggregate"]):
        bb = self.backbone
        outs = (bb.out4, bb.out3, bb.out2)
        stages = [dict(q4=None, cq4=None, s4=[]) for _ in outs]
        for v, img in enumerate(views):
            for st, conv, agg, m in zip(stages, outs, aggregates, self.trunk(img)):
                a, b = ops.fpn_out_prepped(m, conv.weight, agg.ngroups, agg.depth_weight[0].conv.weight, v == 0)
                if v == 0:
                    st["q4"], st["cq4"] = a, b
                else:
                    st["s4"].append(a)
        return [PreppedFeatures(st["q4"], st["cq4"], torch.stack(st["s4"], 0)) for st in stages]


def homo_warping(src_fea: Tensor, src_proj: Tensor, ref_proj: Tensor, depth_hypos: Tensor) -> Tensor:
    """net/unit/base.py:85-126: warp one source feature map into the reference frustum, (B,C,D,H,W)."""
    return ops.homo_warp(src_fea, src_proj, ref_proj, depth_hypos)


def homo_aggregate_by_variance(features: Sequence[Tensor], ref_proj: Tensor, src_projs: Sequence[Tensor],
                               depth_hypos: Tensor) -> Tensor:
    """net/unit/homoaggregate.py:49-69: channel-softmax of the warped sources + cross-view variance."""
    return ops.variance_volume(list(features), ref_proj, list(src_projs), depth_hypos)


def depth_regression(prob_volume: Tensor, depth_hypos: Tensor) -> Tensor:
    """net/unit/regress.py:5-7: sum_d prob * hypothesis -> (B,H,W)."""
    return ops.depth_regression(prob_volume, depth_hypos)


def confidence_regress(prob_volume: Tensor, last_confidence=None, n: int = 4, pad=(0, 0, 0, 0, 1, 2)) -> Tensor:
    """net/unit/regress.py:9-25.  The bicubic blend with `last_confidence` (never taken by
    core.py:75) is done with torch on the op's output, as the reference does."""
    if tuple(pad[:4]) != (0, 0, 0, 0):
        raise ValueError("confidence_regress: only the depth axis may be padded (pad[:4] must be 0)")
    conf = ops.confidence(prob_volume, n, int(pad[4]), int(pad[5]), 1)
    if last_confidence is not None:
        last = torch.nn.functional.interpolate(last_confidence.unsqueeze(1), scale_factor=2, mode="bicubic").squeeze(1)
        conf = 0.8 * last + 0.2 * conf
    return conf


def softmax_regress(logits: Tensor, depth_hypos: Tensor, want_confidence: bool = False, upsample: int = 2):
    """Fused tail for a regulariser that hands over logits (regular.py:67-69,130-133 without its last
    line): returns (prob_volume, depth[, confidence at `upsample` x resolution])."""
    prob, depth, conf = ops.softmax_regress(logits, depth_hypos, True, want_confidence, 4, 1, 2, upsample)
    return (prob, depth, conf) if want_confidence else (prob, depth)


class HyposByFit(nn.Module):
    """Depth hypotheses of a stage, mirroring net/unit/depthhypos.py:10-76 (same constructor and forward).

    depth is None (stage 0): `ndepths` uniform hypotheses over depth_range, (B,ndepths,1,1) -- B*ndepths numbers,
    computed with the same two torch ops as the reference (:31-38).  Otherwise: per-pixel curve fit of the previous
    stage's probability volume ("gauss1", "laplace" or the unwired "gauss0"), x2 upsampling of the fitted scale and of the depth, search
    range from prob_thresh, the reference's clamps, `ndepths` hypotheses per pixel -- two kernels of libmdf_b200.so
    instead of ~40 ATen launches, a batched 3x3 torch.inverse and Python loops over planes and batch items.
    No gradient flows through the reference's version either (depthhypos.py:40 is under no_grad).
    """

    def __init__(self, ndepths: int = 16, curve_calss: str = "gauss1", prob_thresh: float = 0.95):
        super().__init__()
        self.ndepths, self.curve_calss, self.prob_thresh = ndepths, curve_calss, torch.tensor(prob_thresh)

    def forward(self, depth, depth_range, prob_volume, depth_hypos, upsample: bool = False):
        B = depth_range.shape[0]
        if depth is None:
            dmin, dmax = depth_range[:, 0].float().view(B, 1), depth_range[:, 1].float().view(B, 1)
            interval = (dmax - dmin) / (self.ndepths - 1)
            steps = torch.arange(0, self.ndepths, device=depth_range.device).reshape(1, -1)
            return (dmin + steps * interval).view(B, self.ndepths, 1, 1)
        if self.curve_calss not in ("gauss0", "gauss1", "laplace"):       # depthhypos.py:42-47 knows no other
            raise NotImplementedError(f"HyposByFit: unknown curve {self.curve_calss!r}")
        with torch.no_grad():
            s = ops.hypos_fit(prob_volume, depth_hypos, depth, self.curve_calss)
            return ops.hypos_generate(depth, s, depth_range, self.curve_calss, float(self.prob_thresh), self.ndepths,
                                      bool(upsample))


def check_geometric_consistency(depth_ref: Tensor, intrinsics_ref: Tensor, extrinsics_ref: Tensor, depth_src: Tensor,
                                intrinsics_src: Tensor, extrinsics_src: Tensor, thre1=4, thre2=1300.0):
    """tools/filter/dynamic_filter_gpu.py:161-182, same signature and return value: (list of the 9 dynamic masks
    (1,H,W) bool, the loosest mask, depth_reprojected (1,H,W) zeroed outside it) for one (reference, source) pair."""
    out = ops.geo_filter(depth_ref, intrinsics_ref, extrinsics_ref, [depth_src], intrinsics_src.reshape(1, 3, 3),
                         extrinsics_src.reshape(1, 4, 4), None, 0.0, 1, float(thre1), float(thre2), per_source=True)
    bits = out["bits"]
    masks = [((bits >> i) & 1).bool() for i in range(9)]
    return masks, masks[-1], out["depth_reprojected"]


def geometric_filter(ref_depth: Tensor, confidence: Tensor, ref_intrinsics: Tensor, ref_extrinsics: Tensor,
                     src_depths: Sequence[Tensor], src_intrinsics: Tensor, src_extrinsics: Tensor,
                     photo_threshold: float = 0.8, nconditions: int = 5, thre1=4, thre2=1300.0):
    """The per-reference-view body of filter() (tools/filter/dynamic_filter_gpu.py:57-100) in one launch: returns
    (depth_est_averaged (H,W), geo_mask, photo_mask, final_mask (H,W) bool)."""
    out = ops.geo_filter(ref_depth, ref_intrinsics, ref_extrinsics, list(src_depths), src_intrinsics, src_extrinsics,
                         confidence, float(photo_threshold), int(nconditions), float(thre1), float(thre2))
    return out["depth_averaged"], out["geo"], out["photo"], out["final"]
