"""mdf_net_b200 -- B200-native (sm_100a) plane-sweep cost-volume path of MDF-Net.

Public surface = the reference's own unit names for this path (net/unit/homoaggregate.py, net/unit/base.py,
net/unit/regress.py, net/unit/depthhypos.py, net/core.py, tools/filter/dynamic_filter_gpu.py), backed by hand-written
CUDA kernels behind a C ABI
(include/mdf_b200.h, mdf_net_b200/libmdf_b200.so).  Importing the package does not need a GPU;
calling an op without the built library or with CPU tensors raises.
"""
from .core import CoreNet
from .units import (FPNHandOff, HyposByFit, PreppedFeatures, VectorAggregate, check_geometric_consistency, confidence_regress, depth_regression,
                    geometric_filter, homo_aggregate_by_variance, homo_warping, softmax_regress)

__all__ = ["VectorAggregate", "homo_warping", "homo_aggregate_by_variance", "depth_regression", "confidence_regress",
           "softmax_regress", "HyposByFit", "CoreNet", "FPNHandOff", "PreppedFeatures", "check_geometric_consistency", "geometric_filter"]
