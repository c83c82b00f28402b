"""Train-mode VectorAggregate (batch-statistics BatchNorm) and the backward pass of the fused op.

Not implemented yet: raising here keeps the contract "no silent PyTorch fallback".
"""


def vector_aggregate_train(module, features, ref_proj, src_projs, depth_hypos):
    raise NotImplementedError(
        "mdf_net_b200.VectorAggregate: training / autograd is not implemented yet; call it in eval mode under "
        "torch.no_grad() (the reference's eval.py:23-24 does)")
