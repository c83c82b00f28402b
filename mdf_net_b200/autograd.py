"""Train-mode VectorAggregate (batch-statistics BatchNorm) and the backward pass of the fused op.

Reference behaviour (net/unit/homoaggregate.py:25-46 under torch autograd, train.py:33-50): gradients flow to
every feature map and to the depth_weight parameters; projections and hypotheses get none (the sampling
grid is built under no_grad, base.py:97).  In train mode `BatchNorm3d(1)` normalises each source view's z with
that view's batch statistics and updates its running statistics once per source view, in view order
(momentum 0.1, unbiased variance).  All compute is in libmdf_b200.so (csrc/mdf_backward.cu).
"""
from __future__ import annotations

from typing import List

import torch

from . import ops


class _VectorAggregateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, n_feats, groups, training, bn_eps, ref_proj, depth_hypos, conv_w, bn_w, bn_b, bn_mean, bn_var,
                fc_w, fc_b, *tensors):
        features, src_projs = list(tensors[:n_feats]), list(tensors[n_feats:])
        out, stats = ops.cost_volume_train(features, ref_proj, src_projs, depth_hypos, conv_w, bn_w, bn_b, bn_mean, bn_var,
                                           bn_eps, fc_w, fc_b, groups, training)
        # the running statistics may be updated in place after this call: keep the values the forward saw
        # (`stats`: the batch statistics per source view go back into the backward, which then skips its own statistics sweep)
        ctx.save_for_backward(ref_proj, depth_hypos, conv_w, bn_w, bn_b, bn_mean.clone(), bn_var.clone(), fc_w, fc_b, out, stats, *tensors)
        ctx.meta = (n_feats, groups, training, bn_eps)
        ctx.mark_non_differentiable(stats)
        return out, stats

    @staticmethod
    def backward(ctx, grad_out, _grad_stats):
        n_feats, groups, training, bn_eps = ctx.meta
        ref_proj, depth_hypos, conv_w, bn_w, bn_b, bn_mean, bn_var, fc_w, fc_b, out, stats = ctx.saved_tensors[:11]
        tensors = ctx.saved_tensors[11:]
        features, src_projs = list(tensors[:n_feats]), list(tensors[n_feats:])
        gfeats, gp = ops.cost_volume_bwd(features, ref_proj, src_projs, depth_hypos, conv_w, bn_w, bn_b, bn_mean, bn_var,
                                         bn_eps, fc_w, fc_b, groups, training, out, grad_out.contiguous(), stats)
        g_conv = gp[4:].reshape(conv_w.shape)
        return (None, None, None, None, None, None, g_conv, gp[0:1].reshape(bn_w.shape), gp[1:2].reshape(bn_b.shape),
                None, None, gp[2:3].reshape(fc_w.shape), gp[3:4].reshape(fc_b.shape), *gfeats, *([None] * len(src_projs)))


def vector_aggregate_train(module, features: List[torch.Tensor], ref_proj, src_projs, depth_hypos):
    cbr, fc = module.depth_weight[0], module.depth_weight[1]
    bn = cbr.bn
    training = bool(module.training and bn.training)
    out, stats = _VectorAggregateFn.apply(len(features), module.ngroups, training, bn.eps, ref_proj, depth_hypos,
                                          cbr.conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                          fc.weight, fc.bias, *features, *src_projs)
    if training and bn.track_running_stats:
        # BatchNorm3d is applied once per source view: V sequential momentum updates, folded into one expression
        with torch.no_grad():
            V = stats.shape[0]
            bn.num_batches_tracked += V
            if bn.momentum is None:      # cumulative moving average
                n0 = (bn.num_batches_tracked - V).to(stats.dtype)
                bn.running_mean.copy_((bn.running_mean * n0 + stats[:, 0].sum()) / (n0 + V))
                bn.running_var.copy_((bn.running_var * n0 + stats[:, 1].sum()) / (n0 + V))
            else:
                m = float(bn.momentum)
                decay = (1.0 - m) ** torch.arange(V - 1, -1, -1, device=stats.device, dtype=stats.dtype)
                bn.running_mean.mul_((1.0 - m) ** V).add_(m * (decay * stats[:, 0]).sum())
                bn.running_var.mul_((1.0 - m) ** V).add_(m * (decay * stats[:, 1]).sum())
    return out
