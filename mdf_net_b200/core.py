"""Drop-in for the reference's `net.core.CoreNet` (net/core.py:4-78): same constructor, same attribute names (hence the
same state-dict keys), same outputs -- with the tail of every stage taken in ONE launch when the injected units allow it.

What the reference does per stage (core.py:45-65) after the cost volume:

    prob_volume = Regular(cost_volume)          # 3-D CNN ... -> self.prob (Conv3d(c0,1,3)) -> F.softmax   regular.py:67-69
    depth       = Depth_regress(prob_volume, depth_hypos)                                                regress.py:5-7
    (next stage) Depth_hypos(depth, depth_range, prob_volume, depth_hypos, upsample=True)                 depthhypos.py:27-76
    (last stage) Confidence_regress(prob_volume) + nearest x2                                             core.py:75-77

Here the regulariser's body still runs as the injected module (PyTorch / cuDNN, out of scope), but it is stopped in front
of its last layer: `ops.prob_head` takes that layer's input and weight and produces depth, the fitted scale the next
stage's hypotheses need, and the confidence, without the logits or the probability volume ever being written
(SURVEY 8f rows 1-2).  Anything that does not fit (other module types, gradients required, D not in {8,24,48}) goes
through the injected units exactly like the reference (`fuse=False` forces that path).  The tail is fused only for the
regularisers and regress functions it reimplements (the reference's RegularNet_3Scales / _4Scales, depth_regression,
confidence_regress, or this package's); a custom module opts in with a class attribute `mdf_fusable_tail = True`
(its forward must end with `self.prob(x).squeeze(1)` -> softmax over D) or the caller passes `fuse="force"`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class _Captured(Exception):
    """Carries the input of a regulariser's last layer out of its forward()."""

    def __init__(self, x):
        self.x = x


def _fusable_prob_layer(reg: nn.Module):
    conv = getattr(reg, "prob", None)
    ok = (isinstance(conv, nn.Conv3d) and conv.out_channels == 1 and conv.bias is None and conv.groups == 1
          and tuple(conv.kernel_size) == (3, 3, 3) and tuple(conv.stride) == (1, 1, 1)
          and tuple(conv.padding) == (1, 1, 1) and tuple(conv.dilation) == (1, 1, 1) and conv.padding_mode == "zeros"
          and conv.in_channels <= 64)
    return conv if ok else None


def regulariser_body(reg: nn.Module, cost_volume: torch.Tensor) -> torch.Tensor:
    """Run `reg.forward` up to (not including) its last layer `reg.prob` and return that layer's input
    (regular.py:58-67 / :121-130).  The module itself is untouched: a forward pre-hook on `reg.prob` captures."""
    def stop(_mod, args):
        raise _Captured(args[0])

    handle = reg.prob.register_forward_pre_hook(stop)
    try:
        reg(cost_volume)
    except _Captured as c:
        return c.x
    finally:
        handle.remove()
    raise RuntimeError("mdfnet_b200: the regulariser never called its `prob` layer")


class CoreNet(nn.Module):
    """net/core.py:4-78 with fused stage tails.  forward(origin_imgs, extrinsics, intrinsics, depth_range) ->
    {"depth", "confidence"} in eval mode, {"depth": [per-stage depths..., refined]} in training mode."""

    def __init__(self, Backbone, Depth_hypos, scale, Homoaggre, Regular, Regress, Refine, fuse: bool = True,
                 fpn_handoff: bool = False):
        super().__init__()
        # fpn_handoff: (inference, opt-in) run the backbone without its 1x1 output convolutions and let the library apply them
        # straight into the cost-volume kernel's input layout (units.FPNHandOff)
        self.fpn_handoff = fpn_handoff
        self.Backbone = Backbone
        self.Depth_hypos = Depth_hypos
        self.scale = scale
        self.Homoaggre = Homoaggre
        self.Regular = Regular
        self.Depth_regress, self.Confidence_regress = Regress
        self.Refine = Refine
        self.fuse = fuse

    # ------------------------------------------------------------------------------------------------
    def _known_units(self, stage: int) -> bool:
        """The fused tail replaces Regular's last two lines, Depth_regress and Confidence_regress: it may only do so when those
        are the functions it reimplements -- the reference's (net/unit/regress.py, net/unit/regular.py) or this package's --
        never a user's own callable that merely has a `.prob` attribute.  `fuse="force"` skips the check."""
        if self.fuse == "force":
            return True
        def known(fn, name):
            return getattr(fn, "__name__", "") == name and getattr(fn, "__module__", "").rsplit(".", 1)[-1] in ("regress", "units")
        reg_ok = type(self.Regular[stage]).__name__ in ("RegularNet_3Scales", "RegularNet_4Scales") or getattr(self.Regular[stage], "mdf_fusable_tail", False)
        return reg_ok and known(self.Depth_regress, "depth_regression") and known(self.Confidence_regress, "confidence_regress")

    def _can_fuse(self, stage: int, cost_volume: torch.Tensor) -> bool:
        if not self.fuse or self.training or not cost_volume.is_cuda or not self._known_units(stage):
            return False
        if torch.is_grad_enabled() and (cost_volume.requires_grad or any(p.requires_grad for p in self.Regular[stage].parameters())):
            return False              # the fused tail has no backward: training / fine-tuning goes through the injected units
        if cost_volume.shape[2] not in (8, 24, 48) or _fusable_prob_layer(self.Regular[stage]) is None:
            return False
        if stage + 1 < len(self.Depth_hypos):
            nxt = self.Depth_hypos[stage + 1]
            if getattr(nxt, "curve_calss", None) not in ("gauss1", "laplace") or not hasattr(nxt, "ndepths"):
                return False
        return True

    def forward(self, origin_imgs, extrinsics, intrinsics, depth_range):
        views = torch.unbind(origin_imgs.float(), 1)
        from .units import FPNHandOff, VectorAggregate
        handoff = (self.fpn_handoff and not self.training and not torch.is_grad_enabled() and origin_imgs.is_cuda
                   and FPNHandOff.supports(self.Backbone) and all(isinstance(h, VectorAggregate) for h in self.Homoaggre))
        if handoff:
            prepped = FPNHandOff(self.Backbone)(views, list(self.Homoaggre))
            features = None
        else:
            features = [self.Backbone(img) for img in views]                               # core.py:42
        nstages = len(self.Depth_hypos)
        depth = depth_hypos = prob_volume = fitted = None
        depths, confidence = [], None
        for stage in range(nstages):
            feature = prepped[stage] if handoff else [f[stage] for f in features]
            ref_proj, src_projs = self.scale(intrinsics, extrinsics, stage)                # core.py:52
            unit = self.Depth_hypos[stage]
            if fitted is not None:
                # the previous tail already fitted the curve: only the x2 upsampling + range + clamps are left
                depth_hypos = ops.hypos_generate(depth, fitted, depth_range, unit.curve_calss, float(unit.prob_thresh),
                                                 int(unit.ndepths), True)
            else:
                depth_hypos = unit(depth, depth_range, prob_volume, depth_hypos, upsample=True)   # core.py:55
            cost_volume = self.Homoaggre[stage](feature, ref_proj, src_projs, depth_hypos)        # core.py:58
            last = stage + 1 == nstages
            if self._can_fuse(stage, cost_volume):
                reg = self.Regular[stage]
                x = regulariser_body(reg, cost_volume)
                curve = "" if last else self.Depth_hypos[stage + 1].curve_calss
                _, _, depth, conf, s = ops.prob_head(x, reg.prob.weight, depth_hypos, curve, False, False, last)
                fitted = None if last else s
                confidence = conf if last else None
                prob_volume = None
            else:
                prob_volume = self.Regular[stage](cost_volume)                             # core.py:61
                depth = self.Depth_regress(prob_volume, depth_hypos)                       # core.py:64
                fitted = None
            depths.append(depth)
        depth = self.Refine(depth, depth_range)                                            # core.py:69
        depths.append(depth)
        if self.training:
            return {"depth": depths}
        if confidence is None:
            confidence = self.Confidence_regress(prob_volume)                              # core.py:75-77
            confidence = torch.nn.functional.interpolate(confidence.unsqueeze(1), size=None, scale_factor=2,
                                                         mode="nearest", align_corners=None).squeeze(1)
        return {"depth": depth, "confidence": confidence}
