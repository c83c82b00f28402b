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
        # `stats`: the batch statistics per source view go back into the backward, which then skips its own statistics sweep.
        # The running statistics are updated in place right after a train-mode forward (the backward does not read them then);
        # they are buffers, not autograd inputs: plain attributes, no version check.
        ctx.save_for_backward(ref_proj, depth_hypos, conv_w, bn_w, bn_b, fc_w, fc_b, out, stats, *tensors)
        ctx.running = (bn_mean, bn_var)
        ctx.meta = (n_feats, groups, training, bn_eps)
        ctx.mark_non_differentiable(stats)
        return out, stats

    @staticmethod
    def backward(ctx, grad_out, _grad_stats):
        n_feats, groups, training, bn_eps = ctx.meta
        ref_proj, depth_hypos, conv_w, bn_w, bn_b, fc_w, fc_b, out, stats = ctx.saved_tensors[:9]
        bn_mean, bn_var = ctx.running
        tensors = ctx.saved_tensors[9:]
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
        # BatchNorm3d is applied once per source view: V sequential momentum updates in view order, one launch
        with torch.no_grad():
            ops.bn_running_update(stats, -1.0 if bn.momentum is None else float(bn.momentum), bn.running_mean, bn.running_var,
                                  bn.num_batches_tracked)
    return out
