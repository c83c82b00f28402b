"""the FPN hand-off (optional fast entry).  The library applies the backbone's bias-free 1x1 output
convolutions (net/unit/backbone.py:43-45, 59-63) itself and writes the cost-volume kernel's input layout; the result must be
the NCHW drop-in path's result (rel-L2 < 1e-6: only the summation order inside the 1x1 convolution differs from cuDNN's),
without the layout pass in the launch list."""
import contextlib
import io

import numpy as np
import pytest
import torch

from mdf_net_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("stage", [0, 1, 2])
def test_prepped_entry_matches_the_nchw_entry(stage):
    import mdf_net_b200 as mdf
    from mdf_net_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    h0, w0, N, B, Cin = 576, 800, 5, 1, 64
    H, W = syn.stage_shapes(h0, w0)[stage]
    C, D, G = syn.STAGE_CHANNELS[stage], syn.STAGE_DEPTHS[stage], syn.STAGE_GROUPS[stage]
    K, E = syn.camera_rig(B, N, h0, w0, seed=21)
    P = syn.projection_matrices(K, E, 2.0 ** (3 - stage))
    gen = torch.Generator(device="cuda").manual_seed(stage)
    # the FPN's merged maps: smooth, O(1); the 1x1 output convolution gives features of std ~ 1.5
    xs = [torch.nn.functional.avg_pool2d(torch.randn((B, Cin, H, W), device="cuda", generator=gen), 3, 1, 1) * 3.0 for _ in range(N)]
    conv = torch.nn.Conv2d(Cin, C, 1, bias=False).cuda()
    with torch.no_grad():
        conv.weight.mul_(1.5 / float(conv(xs[0]).std()))
    hyp = cu(syn.uniform_hypos(B, D) if stage == 0 else syn.scene_hypos(B, D, H, W, seed=21))
    agg = mdf.VectorAggregate(G).cuda().eval()
    p = syn.depth_weight_params(G, seed=30 + stage)
    with torch.no_grad():
        dw = agg.depth_weight
        dw[0].conv.weight.copy_(cu(p["cw"]).view(1, G, 1, 1, 1))
        dw[0].bn.weight.fill_(float(p["bn_weight"])); dw[0].bn.bias.fill_(float(p["bn_bias"]))
        dw[0].bn.running_mean.fill_(float(p["bn_mean"])); dw[0].bn.running_var.fill_(float(p["bn_var"]))
        dw[1].weight.fill_(float(p["fc_weight"])); dw[1].bias.fill_(float(p["fc_bias"]))
        ref_proj, src_projs = cu(P[:, 0]), [cu(P[:, v]) for v in range(1, N)]
        feats = [conv(x) for x in xs]                                         # backbone.py:59 / :61 / :63 through cuDNN
        ops.reset_launch_count()
        want = agg(feats, ref_proj, src_projs, hyp)
        n_nchw = ops.launch_count()
        q4, cq4 = ops.fpn_out_prepped(xs[0], conv.weight, G, dw[0].conv.weight, True)
        s4 = torch.stack([ops.fpn_out_prepped(x, conv.weight, G, dw[0].conv.weight, False)[0] for x in xs[1:]], 0)
        ops.reset_launch_count()
        got = agg(mdf.PreppedFeatures(q4, cq4, s4), ref_proj, src_projs, hyp)
        n_prepped = ops.launch_count()
    assert got.shape == want.shape == (B, G, D, H, W)
    r = float(torch.linalg.vector_norm((got - want).double()) / torch.linalg.vector_norm(want.double()))
    assert r < 1e-6, f"stage {stage}: prepped vs NCHW entry rel-L2 {r:.3g}"
    assert (n_nchw, n_prepped) == (3, 2)            # setup + layout pass + hot kernel  vs  setup + hot kernel
    # the maps themselves are what the layout pass would have written from the NCHW features
    d = (feats[1][:, 1::2] - feats[1][:, 0::2]) * 1.4426950408889634                                  # (B,G,H,W)
    planar = d.view(B, G // 4, 4, H, W).permute(0, 1, 3, 4, 2)
    assert float((s4[0] - planar).abs().max()) < 2e-5 * float(planar.abs().max())


def test_corenet_with_the_hand_off_matches_the_plain_drop_in():
    """mdf.CoreNet(fpn_handoff=True) on the reference's own FPN_4Scales (oracle/_ref) against the same model without it."""
    import mdf_net_b200 as mdf
    from oracle import ref_bench, ref_install
    if not ref_install.available():
        pytest.fail("oracle/_ref is missing on this box (python -m oracle.ref_install)")
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with contextlib.redirect_stdout(io.StringIO()):
        r = ref_install.modules()
        torch.manual_seed(7)
        ndepths, ngroups, curves, thresh = (48, 24, 8), (32, 16, 8), [None, "gauss1", "laplace"], (0.0, 0.95, 1e-5)
        parts = lambda: (r.backbone.FPN_4Scales((8, 16, 32, 64)),
                         torch.nn.ModuleList([mdf.HyposByFit(ndepths[i], curves[i], thresh[i]) for i in range(3)]), r.scale.scale_cam,
                         torch.nn.ModuleList([mdf.VectorAggregate(g) for g in ngroups]),
                         torch.nn.ModuleList([r.regular.RegularNet_3Scales(32), r.regular.RegularNet_4Scales(16), r.regular.RegularNet_4Scales(8)]),
                         [mdf.depth_regression, mdf.confidence_regress], r.refine.RefineNet2())
        plain = mdf.CoreNet(*parts())
        fast = mdf.CoreNet(*parts(), fpn_handoff=True)
    assert mdf.FPNHandOff.supports(plain.Backbone)
    h0, w0, N = 512, 640, 3
    K, E = syn.camera_rig(1, N, h0, w0, seed=5)
    imgs = cu(np.random.default_rng(5).random((1, N, 3, h0, w0), dtype=np.float32))
    plain = ref_bench.randomise_weights(plain.cuda(), imgs[:, 0]).eval()
    fast.load_state_dict(plain.state_dict(), strict=True)
    fast = fast.cuda().eval()
    args = (imgs, cu(E), cu(K), cu(np.array([[425.0, 935.0]], np.float32)))
    from mdf_net_b200 import ops
    with torch.no_grad():
        ops.reset_launch_count(); a = plain(*args); n_plain = ops.launch_count()
        ops.reset_launch_count(); b = fast(*args); n_fast = ops.launch_count()
    err = (a["depth"] - b["depth"]).abs()
    assert float((err < 0.5).float().mean()) >= 0.9999 and float(err.median()) < 1e-2, (float(err.median()), float(err.max()))
    assert float(((a["confidence"] > 0.8) == (b["confidence"] > 0.8)).float().mean()) >= 0.9995
    # 3 stages: the layout pass is gone (-3), the 1x1 output convolutions arrive as 3 stages x N views launches of the library
    assert n_fast == n_plain - 3 + 3 * N
