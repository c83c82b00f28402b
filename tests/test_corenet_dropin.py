"""`mdf_net_b200.CoreNet`, the drop-in for net/core.py:4-78.

CPU (build container only, needs the read-only reference checkout): wired with the reference's OWN units and fuse=False
it must reproduce `config.model` bit for bit -- the orchestration is the reference's.
GPU: with this package's units and a small stand-in regulariser (the 3-D CNN is out of scope; what matters is a module
whose last layer is `self.prob = Conv3d(c0, 1, 3, padding=1, bias=False)` followed by softmax, regular.py:43,67-69), the
fused stage tails must give what the unit-by-unit path gives."""
import os
import sys

import numpy as np
import pytest

from conftest import assert_confidence_decisions
import torch
import torch.nn as nn
import torch.nn.functional as F

from mdf_net_b200 import synthetic as syn

REF = os.environ.get("MDF_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "net")), reason="reference checkout not present (GPU box)")
def test_same_orchestration_as_the_reference_corenet():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    try:
        import config                                   # builds config.model (config.py:186-218)
    finally:
        sys.path.remove(REF)
    import mdf_net_b200 as mdf
    ref = config.model.eval()
    mine = mdf.CoreNet(ref.Backbone, ref.Depth_hypos, ref.scale, ref.Homoaggre, ref.Regular,
                       [ref.Depth_regress, ref.Confidence_regress], ref.Refine, fuse=False).eval()
    assert sorted(mine.state_dict().keys()) == sorted(ref.state_dict().keys())
    B, N, H0, W0 = 1, 3, 64, 64
    rng = np.random.default_rng(5)
    imgs = torch.from_numpy(rng.uniform(0, 1, (B, N, 3, H0, W0)).astype(np.float32))
    K, E = syn.camera_rig(B, N, H0, W0, seed=5)
    dr = torch.tensor([[425.0, 935.0]])
    with torch.no_grad():
        a = ref(imgs, torch.from_numpy(E), torch.from_numpy(K), dr)
        b = mine(imgs, torch.from_numpy(E), torch.from_numpy(K), dr)
    assert torch.equal(a["depth"], b["depth"]) and torch.equal(a["confidence"], b["confidence"])
    ref.train(); mine.train()
    with torch.no_grad():
        a = ref(imgs, torch.from_numpy(E), torch.from_numpy(K), dr)
        b = mine(imgs, torch.from_numpy(E), torch.from_numpy(K), dr)
    assert len(a["depth"]) == len(b["depth"]) == 4


class _Backbone(nn.Module):
    """Stand-in feature pyramid: (B,3,H,W) -> features at 1/8, 1/4, 1/2 with 64 / 32 / 16 channels."""

    def __init__(self):
        super().__init__()
        self.heads = nn.ModuleList([nn.Conv2d(3, c, 3, padding=1) for c in (64, 32, 16)])

    def forward(self, img):
        return [h(F.avg_pool2d(img, k)) * 3.0 for h, k in zip(self.heads, (8, 4, 2))]


class _Regular(nn.Module):
    """Stand-in regulariser with the reference's tail: ... -> self.prob -> squeeze -> softmax (regular.py:43,67-69)."""
    mdf_fusable_tail = True       # opts in to the fused tail (custom modules are not fused unless they say so)

    def __init__(self, in_chs, c0):
        super().__init__()
        self.body = nn.Conv3d(in_chs, c0, 3, padding=1)
        self.prob = nn.Conv3d(c0, 1, 3, stride=1, padding=1, bias=False)

    def forward(self, x):
        x = F.relu(self.body((x - 0.5) * 8.0))
        x = self.prob(x).squeeze(1)
        return F.softmax(x, dim=1)


class _Refine(nn.Module):
    def forward(self, depth, depth_range):
        return F.interpolate(depth.unsqueeze(1), scale_factor=2, mode="bilinear").squeeze(1)


def _scale_cam(intrinsics, extrinsics, stage):
    """scale.py:4-20 restated (pinned by tests/golden/scale_cam.npz in test_oracle_golden.py)."""
    K = intrinsics.clone()
    K[:, :, :2, :] = K[:, :, :2, :] / (2 ** (3 - stage))
    P = extrinsics.clone()
    P[:, :, :3, :4] = torch.matmul(K, extrinsics[:, :, :3, :4])
    views = torch.unbind(P, 1)
    return views[0], views[1:]


@pytest.mark.gpu
def test_fused_stage_tails_match_the_unit_by_unit_path():
    import mdf_net_b200 as mdf
    from mdf_net_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(3)
    dev = torch.device("cuda")
    hyp = nn.ModuleList([mdf.HyposByFit(d, c, t) for d, c, t in zip((48, 24, 8), (None, "gauss1", "laplace"), (0.0, 0.95, 1e-5))])
    agg = nn.ModuleList([mdf.VectorAggregate(g) for g in (32, 16, 8)])
    reg = nn.ModuleList([_Regular(32, 16), _Regular(16, 8), _Regular(8, 8)])
    with torch.no_grad():
        for r in reg:
            r.prob.weight.mul_(6.0)
    args = (_Backbone(), hyp, _scale_cam, agg, reg, [mdf.depth_regression, mdf.confidence_regress], _Refine())
    fused = mdf.CoreNet(*args, fuse=True).to(dev).eval()
    plain = mdf.CoreNet(*args, fuse=False).to(dev).eval()
    assert sorted(fused.state_dict().keys()) == sorted(plain.state_dict().keys())
    B, N, H0, W0 = 1, 4, 256, 320
    rng = np.random.default_rng(11)
    imgs = torch.from_numpy(rng.uniform(0, 1, (B, N, 3, H0, W0)).astype(np.float32)).to(dev)
    K, E = syn.camera_rig(B, N, H0, W0, seed=11)
    K, E = torch.from_numpy(K).to(dev), torch.from_numpy(E).to(dev)
    dr = torch.tensor([[425.0, 935.0]], device=dev)
    with torch.no_grad():
        ops.reset_launch_count()
        a = fused(imgs, E, K, dr)
        n_fused = ops.launch_count()
        ops.reset_launch_count()
        seen = []
        hook = plain.Regular[2].register_forward_hook(lambda m, i, o: seen.append(o))
        b = plain(imgs, E, K, dr)
        hook.remove()
        n_plain = ops.launch_count()
    interval = (935.0 - 425.0) / 47.0
    # the convolution's summation order differs from cuDNN's (1e-6 of the logits); three chained stages
    assert (a["depth"] - b["depth"]).abs().max().item() < 2e-3 * interval
    # confidence: exact count of differing pixels, each one a truncation flip of the expected index (conftest)
    assert_confidence_decisions(a["confidence"].cpu().numpy(), b["confidence"].cpu().numpy(), seen[0].cpu().numpy(), "fused vs unit by unit",
                                index_noise=1e-3)
    assert a["confidence"].shape == (B, H0, W0) and a["depth"].shape == (B, H0, W0)
    assert n_fused < n_plain                    # 3 x (cost volume + tail) + 2 hypothesis launches vs the split units
    # training mode -> the reference's training output (core.py:72-73), through the injected units
    fused.train()
    with torch.no_grad():
        out = fused(imgs, E, K, dr)
    assert isinstance(out["depth"], list) and len(out["depth"]) == 4


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "net")), reason="reference checkout not present (GPU box)")
def test_regulariser_body_stops_the_reference_modules_in_front_of_prob():
    """`regulariser_body` on the reference's own RegularNet_3Scales / RegularNet_4Scales: what it returns is exactly the
    tensor their forward() feeds to `self.prob` (conv + squeeze + softmax of it reproduce the module's output), the
    module is left untouched (no hook stays behind), and `_fusable_prob_layer` accepts both."""
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    try:
        from net.unit.regular import RegularNet_3Scales, RegularNet_4Scales
    finally:
        sys.path.remove(REF)
    from mdf_net_b200.core import _fusable_prob_layer, regulariser_body
    torch.manual_seed(0)
    for net, shape in ((RegularNet_3Scales(32).eval(), (1, 32, 8, 8, 12)), (RegularNet_4Scales(8).eval(), (2, 8, 8, 16, 16))):
        cv = torch.rand(shape)
        with torch.no_grad():
            x = regulariser_body(net, cv)
            want = net(cv)
            got = F.softmax(net.prob(x).squeeze(1), dim=1)
        assert x.shape == (shape[0], net.prob.in_channels) + shape[2:]
        assert torch.equal(got, want)
        assert len(net.prob._forward_pre_hooks) == 0
        assert _fusable_prob_layer(net) is net.prob


def test_fused_tail_only_replaces_the_units_it_reimplements():
    """A user's own regress callable, or a regulariser that merely has a `.prob` layer, is never bypassed silently."""
    import mdf_net_b200 as mdf

    class Plain(nn.Module):                     # has .prob but did not opt in
        def __init__(self):
            super().__init__()
            self.prob = nn.Conv3d(8, 1, 3, padding=1, bias=False)

    def my_depth(prob_volume, depth_hypos):      # a custom regression with another name
        return (prob_volume * depth_hypos).sum(1)

    hyp = nn.ModuleList([mdf.HyposByFit(8, None, 0.0)])
    args = lambda reg, regress: (nn.Identity(), hyp, None, nn.ModuleList([mdf.VectorAggregate(8)]), nn.ModuleList([reg]), regress, nn.Identity())
    ok = mdf.CoreNet(*args(_Regular(8, 8), [mdf.depth_regression, mdf.confidence_regress]))
    assert ok._known_units(0)
    assert not mdf.CoreNet(*args(Plain(), [mdf.depth_regression, mdf.confidence_regress]))._known_units(0)
    assert not mdf.CoreNet(*args(_Regular(8, 8), [my_depth, mdf.confidence_regress]))._known_units(0)
    assert mdf.CoreNet(*args(Plain(), [my_depth, mdf.confidence_regress]), fuse="force")._known_units(0)
