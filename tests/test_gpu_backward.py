"""GPU parity of the train-mode forward and the backward pass (BASELINE.json configs[4]).

References: (1) gradients of the unmodified reference under torch autograd (tests/golden/vecagg_grad_*.npz),
(2) tests/torch_ref.py -- a plain-PyTorch restatement pinned on (1) by the CPU suite -- evaluated on the GPU
in float64 (mid sizes and the BlendedMVS train shape, batch 8).  Tolerance: cost volume rel-L2 < 1e-5; gradients as close
to float64 autograd as float32 autograd of the same formulae is, and rel-L2 < 1e-4 where float32 autograd is that good.
"""
import numpy as np
import pytest
import torch

import torch_ref
from conftest import load_golden, rel_l2
from mdf_net_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def cu(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype)


def make_module(G, p):
    import mdf_net_b200 as mdf
    m = mdf.VectorAggregate(G).cuda()
    with torch.no_grad():
        dw = m.depth_weight
        dw[0].conv.weight.copy_(cu(p["cw"]).view(1, G, 1, 1, 1))
        dw[0].bn.weight.fill_(float(p["bn_weight"])); dw[0].bn.bias.fill_(float(p["bn_bias"]))
        dw[0].bn.running_mean.fill_(float(p["bn_mean"])); dw[0].bn.running_var.fill_(float(p["bn_var"]))
        dw[1].weight.fill_(float(p["fc_weight"])); dw[1].bias.fill_(float(p["fc_bias"]))
    return m


def run_module(m, feats, ref_proj, src_projs, hyp, gout):
    fs = [cu(f).requires_grad_(True) for f in feats]
    cv = m(fs, cu(ref_proj), [cu(s) for s in src_projs], cu(hyp))
    cv.backward(cu(gout))
    dw = m.depth_weight
    return dict(cv=cv.detach().cpu().numpy(), gf=np.stack([f.grad.cpu().numpy() for f in fs]),
                gcw=dw[0].conv.weight.grad.cpu().numpy().reshape(-1),
                gbn=np.array([dw[0].bn.weight.grad.item(), dw[0].bn.bias.grad.item()]),
                gfc=np.array([dw[1].weight.grad.item(), dw[1].bias.grad.item()]))


@pytest.mark.parametrize("mode", ["eval", "train"])
@pytest.mark.parametrize("name", ["vecagg_grad_s0", "vecagg_grad_s2"])
def test_gradients_golden(name, mode):
    z = load_golden(name)
    p = {k[2:]: z[k] for k in z.files if k.startswith("p_")}
    m = make_module(int(z["groups"]), p)
    m.train(mode == "train")
    r = run_module(m, z["features"], z["ref_proj"], z["src_projs"], z["depth_hypos"], z["grad_out"])
    assert rel_l2(r["cv"], z[f"{mode}_cost_volume"]) < 1e-5
    assert rel_l2(r["gf"], z[f"{mode}_grad_features"]) < 1e-4
    assert rel_l2(r["gcw"], z[f"{mode}_grad_cw"]) < 2e-4
    assert np.allclose(r["gbn"], z[f"{mode}_grad_bn"], rtol=5e-4, atol=2e-5)
    assert np.allclose(r["gfc"], z[f"{mode}_grad_fc"], rtol=5e-4, atol=2e-5)
    bn = m.depth_weight[0].bn
    if mode == "train":    # running statistics: one momentum update per source view, in view order
        got = [bn.running_mean.item(), bn.running_var.item(), float(bn.num_batches_tracked.item())]
        assert np.allclose(got, z["train_running"], rtol=1e-5)
    else:
        assert bn.running_mean.item() == pytest.approx(float(p["bn_mean"])) and int(bn.num_batches_tracked.item()) == 0


def reference_grads(feats, ref_proj, src_projs, hyp, p, G, gout, training, dtype):
    fs = [cu(f, dtype).requires_grad_(True) for f in feats]
    P = {k: cu(np.asarray(p[k], np.float32).reshape(-1), dtype).requires_grad_(True)
         for k in ("cw", "bn_weight", "bn_bias", "fc_weight", "fc_bias")}
    cv, _ = torch_ref.vector_aggregate(fs, cu(ref_proj, dtype), [cu(s, dtype) for s in src_projs], cu(hyp, dtype),
                                       P["cw"], P["bn_weight"], P["bn_bias"], float(p["bn_mean"]), float(p["bn_var"]),
                                       float(p["bn_eps"]), P["fc_weight"], P["fc_bias"], G, training=training)
    cv.backward(cu(gout, dtype))
    return dict(cv=cv.detach().cpu().numpy(), gf=np.stack([f.grad.cpu().numpy() for f in fs]),
                gcw=P["cw"].grad.cpu().numpy(), gbn=np.array([P["bn_weight"].grad.item(), P["bn_bias"].grad.item()]),
                gfc=np.array([P["fc_weight"].grad.item(), P["fc_bias"].grad.item()]))


def case(stage, h0, w0, nviews, batch, seed):
    H, W = syn.stage_shapes(h0, w0)[stage]
    C, D, G = syn.STAGE_CHANNELS[stage], syn.STAGE_DEPTHS[stage], syn.STAGE_GROUPS[stage]
    K, E = syn.camera_rig(batch, nviews, h0, w0, seed=seed)
    P = syn.projection_matrices(K, E, level_div=2.0 ** (3 - stage))
    feats = syn.smooth_features(batch, nviews, C, H, W, seed=seed + stage)
    hyp = syn.uniform_hypos(batch, D) if stage == 0 else syn.scene_hypos(batch, D, H, W, seed=seed + stage)
    gout = np.random.default_rng(seed).standard_normal((batch, G, D, H, W)).astype(np.float32)
    return feats, P[:, 0], [P[:, v] for v in range(1, nviews)], hyp, syn.depth_weight_params(G, seed=seed + stage), G, gout


def compare(r, ref, ref32, what):
    """The CUDA gradients must be as close to float64 autograd as float32 autograd of the same formulae is
    (train-mode BatchNorm subtracts batch means of the gradient: float32 cancellation costs ~1e-3 there,
    tools/diag_backward.py), and never worse than 1e-4 where float32 autograd itself is that good."""
    def check(key, floor):
        mine, theirs = rel_l2(r[key], ref[key]), rel_l2(ref32[key], ref[key])
        assert mine < max(floor, 1.5 * theirs), f"{what} {key}: cuda {mine:.2e}, float32 autograd {theirs:.2e}"
    assert rel_l2(r["cv"], ref["cv"]) < 1e-5, what
    check("gf", 1e-4)
    if np.abs(ref["gcw"]).max() > 0:
        check("gcw", 2e-4)
    for key in ("gbn", "gfc"):
        tol = np.maximum(3.0 * np.abs(ref32[key] - ref[key]), 1e-4 + 5e-4 * np.abs(ref[key]))
        assert (np.abs(r[key] - ref[key]) <= tol).all(), (what, key, r[key], ref[key], ref32[key])


@pytest.mark.parametrize("training", [False, True])
@pytest.mark.parametrize("stage", [0, 1, 2])
def test_gradients_vs_float64_autograd(stage, training):
    feats, ref_proj, src_projs, hyp, p, G, gout = case(stage, 256, 320, 4, 2, seed=500)
    m = make_module(G, p)
    m.train(training)
    r = run_module(m, feats, ref_proj, src_projs, hyp, gout)
    ref = reference_grads(feats, ref_proj, src_projs, hyp, p, G, gout, training, torch.float64)
    ref32 = reference_grads(feats, ref_proj, src_projs, hyp, p, G, gout, training, torch.float32)
    compare(r, ref, ref32, f"stage {stage} training={training}")


@pytest.mark.parametrize("stage", [0, 1, 2])
def test_blendedmvs_train_shape_batch8(stage):
    """BASELINE.json configs[4]: 768x576, N=5, batch 8, train-mode forward + backward, all three stages, against FLOAT64
    autograd of the restatement with the mid-size criterion: the CUDA gradients are as close to float64 as float32
    autograd of the same formulae is (the float64 graph of the whole batch fits the 180 GB of a B200)."""
    feats, ref_proj, src_projs, hyp, p, G, gout = case(stage, 576, 768, 5, 8, seed=600)
    m = make_module(G, p)
    m.train(True)
    r = run_module(m, feats, ref_proj, src_projs, hyp, gout)
    ref = reference_grads(feats, ref_proj, src_projs, hyp, p, G, gout, True, torch.float64)
    torch.cuda.empty_cache()
    ref32 = reference_grads(feats, ref_proj, src_projs, hyp, p, G, gout, True, torch.float32)
    torch.cuda.empty_cache()
    compare(r, ref, ref32, f"BlendedMVS train shape, stage {stage}")


def test_backward_is_linear_and_eval_train_paths_agree():
    from mdf_net_b200 import ops
    feats, ref_proj, src_projs, hyp, p, G, gout = case(2, 256, 320, 3, 1, seed=700)
    m = make_module(G, p).eval()
    a = run_module(m, feats, ref_proj, src_projs, hyp, gout)
    m.zero_grad()
    b = run_module(m, feats, ref_proj, src_projs, hyp, 2.0 * gout)
    assert rel_l2(b["gf"], 2.0 * a["gf"]) < 1e-5
    # eval mode under autograd and the tuned no_grad kernel compute the same volume
    with torch.no_grad():
        fast = m([cu(f) for f in feats], cu(ref_proj), [cu(s) for s in src_projs], cu(hyp)).cpu().numpy()
    assert rel_l2(fast, a["cv"]) < 1e-5


@pytest.mark.parametrize("momentum", [0.1, 0.37, None])
def test_running_statistics_update_is_v_sequential_batchnorm_updates(momentum):
    """mdf_bn_running_update = what BatchNorm3d does to its buffers when the module is applied once per source view
    (base.py:50-68 inside homoaggregate.py:40): V momentum updates in view order, or the cumulative average when momentum is None."""
    from mdf_net_b200 import ops
    V = 6
    rng = np.random.default_rng(11)
    stats = rng.random((V, 2)).astype(np.float32) + 0.1
    bn = torch.nn.BatchNorm3d(1, momentum=momentum).cuda().train()
    with torch.no_grad():
        bn.running_mean.fill_(0.3); bn.running_var.fill_(1.7); bn.num_batches_tracked.fill_(5)
    rm, rv, nbt = bn.running_mean.clone(), bn.running_var.clone(), bn.num_batches_tracked.clone()
    ops.bn_running_update(cu(stats), -1.0 if momentum is None else momentum, rm, rv, nbt)
    # the reference: feed BatchNorm3d inputs whose batch statistics are exactly stats[v]
    for v in range(V):
        n = 4096
        x = torch.randn(n, dtype=torch.float64, device="cuda")
        x = (x - x.mean()) / x.std(unbiased=True) * float(np.sqrt(stats[v, 1])) + float(stats[v, 0])
        bn(x.float().view(n, 1, 1, 1, 1))
    assert int(nbt.item()) == int(bn.num_batches_tracked.item()) == 5 + V
    assert rm.item() == pytest.approx(bn.running_mean.item(), rel=2e-5)
    assert rv.item() == pytest.approx(bn.running_var.item(), rel=2e-5)
    # argument validation
    with pytest.raises(RuntimeError):
        ops.bn_running_update(cu(stats[:, :1]), 0.1, rm, rv, nbt)
    with pytest.raises(RuntimeError):
        ops.bn_running_update(cu(stats), 0.1, rm, rv, nbt.float())
