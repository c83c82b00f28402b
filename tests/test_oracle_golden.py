"""Pin the CPU oracle (oracle/mdf_oracle.c) on outputs of the unmodified reference.

The fixtures under tests/golden/ were produced by tests/golden/make_golden.py, which imports
the reference's own homo_warping / VectorAggregate / regress functions (torch CPU).
Tolerances: the coordinate + bilinear restatement is bit-exact once the reference's own 4x4
`src_proj @ inverse(ref_proj)` is supplied; with the oracle's own 4x4 LU the projection differs
from MKL's by 1-2 ulp, which moves samples by ~1e-5 px (rel-L2 ~1e-6 on warped features).
"""
import numpy as np
import pytest

from oracle import c_oracle as co
from conftest import assert_confidence_decisions, load_golden, rel_l2, max_abs_over_max


def _rt_from_proj(proj):
    B = proj.shape[0]
    return np.concatenate([proj[:, :3, :3].reshape(B, 9), proj[:, :3, 3]], 1)


@pytest.mark.parametrize("name", ["warp_uniform", "warp_pixel"])
def test_warp_bit_exact_given_reference_projection(name):
    z = load_golden(name)
    for v in range(z["src_projs"].shape[0]):
        w = co.homo_warp(z["src_fea"], z["depth_hypos"], rot_trans=_rt_from_proj(z["proj"][v]))
        assert np.array_equal(w, z["warped"][v]), "coordinate/bilinear restatement must be bit exact"


@pytest.mark.parametrize("name", ["warp_uniform", "warp_pixel"])
def test_compose_proj_and_warp(name):
    z = load_golden(name)
    for v in range(z["src_projs"].shape[0]):
        rt = co.compose_proj(z["src_projs"][v], z["ref_proj"])
        ref = _rt_from_proj(z["proj"][v])
        assert np.abs(rt - ref).max() <= 4e-7 * np.abs(ref).max()
        w = co.homo_warp(z["src_fea"], z["depth_hypos"], src_proj=z["src_projs"][v], ref_proj=z["ref_proj"])
        assert rel_l2(w, z["warped"][v]) < 1e-5


def test_warp_edge_cases():
    """z<=0, tiny, huge depths follow the reference exactly; non-finite hypotheses give zeros
    (the reference's CUDA grid_sample maps NaN to -100, GridSampler.cuh:140-147; its CPU kernel
    returns NaN there -- documented divergence)."""
    z = load_golden("warp_edge")
    w = co.homo_warp(z["src_fea"], z["depth_hypos"], src_proj=z["src_proj"], ref_proj=z["ref_proj"])
    ref = z["warped"]
    finite = np.isfinite(z["depth_hypos"].ravel())
    assert rel_l2(w[:, :, finite], ref[:, :, finite]) < 1e-5
    assert np.all((w[:, :, finite] != 0) == (ref[:, :, finite] != 0))
    assert np.all(np.isnan(ref[:, :, ~finite])) and np.all(w[:, :, ~finite] == 0)
    # z == 0 and depth 1e-3 put every sample far outside: all zeros in both
    assert np.all(ref[:, :, 2:4] == 0) and np.all(w[:, :, 2:4] == 0)
    wi = co.homo_warp(z["src_fea"], z["depth_hypos"][:, :1], src_proj=z["ref_proj"], ref_proj=z["ref_proj"])
    assert rel_l2(wi, z["warped_identity"]) < 1e-5
    # identity homography is NOT the identity map: half-pixel shift of the align_corners mismatch
    assert rel_l2(wi[:, :, 0], z["src_fea"]) > 1e-2


def _vecagg(z, prec="f32"):
    p = {k[2:]: z[k] for k in z.files if k.startswith("p_")}
    return co.vector_aggregate(list(z["features"]), z["depth_hypos"], p, int(z["groups"]),
                               ref_proj=z["ref_proj"], src_projs=list(z["src_projs"]), prec=prec)


@pytest.mark.parametrize("name", ["vecagg_s0", "vecagg_s1", "vecagg_s2", "vecagg_cpg4", "vecagg_n2"])
def test_vector_aggregate(name):
    z = load_golden(name)
    out = _vecagg(z)
    ref = z["cost_volume"]
    assert out.shape == ref.shape
    assert rel_l2(out, ref) < 1e-6
    assert max_abs_over_max(out, ref) < 1e-5
    # noise floor: the oracle is as close to a float64 evaluation as the reference itself
    truth = _vecagg(z, "f64")
    assert rel_l2(out, truth) < 3 * max(rel_l2(ref, truth), 5e-8)


@pytest.mark.parametrize("name", ["varagg_uniform", "varagg_pixel"])
def test_variance_aggregate(name):
    z = load_golden(name)
    out = co.variance_aggregate(list(z["features"]), z["depth_hypos"], ref_proj=z["ref_proj"],
                                src_projs=list(z["src_projs"]))
    assert rel_l2(out, z["cost_volume"]) < 1e-6
    assert max_abs_over_max(out, z["cost_volume"]) < 1e-5


@pytest.mark.parametrize("name", ["head_d48", "head_d24", "head_d8"])
def test_head(name):
    z = load_golden(name)
    assert np.abs(co.softmax_depth(z["logits"]) - z["prob"]).max() < 3e-7
    interval = (935.0 - 425.0) / 47.0   # stage-0 hypothesis interval (depthhypos.py:33)
    for kind in ("uniform", "pixel"):
        d = co.depth_regression(z["prob"], z["hypos_" + kind])
        assert np.abs(d - z["depth_" + kind]).max() < 1e-3 * interval
    assert np.array_equal(co.confidence_regress(z["prob"]), z["confidence"])
    assert np.array_equal(co.confidence_regress(z["prob"], upsample=2), z["confidence_up"])


def test_head_known_answers():
    z = load_golden("head_known")
    assert np.array_equal(co.confidence_regress(z["onehot"]), z["onehot_conf"])
    assert np.all(z["onehot_conf"] == 1.0)          # window k-1..k+2 always holds the one-hot bin
    assert np.array_equal(co.confidence_regress(z["uniform"]), z["uniform_conf"])
    assert np.all(z["uniform_conf"] == 0.5)         # E[idx]=3.5 -> 3, window of 4 out of 8


def test_degenerate_features_give_half():
    """All-equal features: every group similarity is 0.5 regardless of the weights."""
    z = load_golden("vecagg_s2")
    feats = [np.full_like(f, 0.37) for f in z["features"]]
    p = {k[2:]: z[k] for k in z.files if k.startswith("p_")}
    out = co.vector_aggregate(feats, z["depth_hypos"], p, int(z["groups"]), ref_proj=z["ref_proj"],
                              src_projs=list(z["src_projs"]))
    assert np.abs(out - 0.5).max() < 1e-6


def test_corenet_stages_teacher_forced():
    """Replay every stage of a whole reference CoreNet forward (captured by hooks)."""
    z = load_golden("corenet_64x64_n3")
    for s in range(3):
        bn, fc = z[f"s{s}_bn"], z[f"s{s}_fc"]
        p = {"cw": z[f"s{s}_cw"], "bn_weight": bn[0], "bn_bias": bn[1], "bn_mean": bn[2], "bn_var": bn[3],
             "bn_eps": bn[4], "fc_weight": fc[0], "fc_bias": fc[1]}
        G = z[f"s{s}_cw"].size
        out = co.vector_aggregate(list(z[f"s{s}_features"]), z[f"s{s}_depth_hypos"], p, G,
                                  ref_proj=z[f"s{s}_ref_proj"], src_projs=list(z[f"s{s}_src_projs"]))
        assert rel_l2(out, z[f"s{s}_cost_volume"]) < 1e-5, s
        prob = co.softmax_depth(z[f"s{s}_logits"])
        assert np.abs(prob - z[f"s{s}_prob"]).max() < 3e-7
        d = co.depth_regression(z[f"s{s}_prob"], z[f"s{s}_depth_hypos"])
        assert np.abs(d - z[f"s{s}_depth"]).max() < 1e-3 * (935.0 - 425.0) / 47.0
    conf = co.confidence_regress(z["s2_prob"], upsample=2)
    assert np.array_equal(conf, z["confidence"])


@pytest.mark.parametrize("name", ["vecagg_grad_s0", "vecagg_grad_s2"])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_torch_restatement_gradients(name, mode):
    """tests/torch_ref.py (the autograd reference of the CUDA backward) reproduces the unmodified reference's
    forward and gradients, eval and train mode, and its batch statistics reproduce the running-stat updates."""
    import torch
    import torch_ref
    z = load_golden(name)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    feats = [T(f).clone().requires_grad_(True) for f in z["features"]]
    P = {k: T(np.asarray(z["p_" + k], np.float32).reshape(-1)).clone().requires_grad_(True)
         for k in ("cw", "bn_weight", "bn_bias", "fc_weight", "fc_bias")}
    cv, stats = torch_ref.vector_aggregate(feats, T(z["ref_proj"]), [T(s) for s in z["src_projs"]], T(z["depth_hypos"]),
                                           P["cw"], P["bn_weight"], P["bn_bias"], float(z["p_bn_mean"]), float(z["p_bn_var"]),
                                           float(z["p_bn_eps"]), P["fc_weight"], P["fc_bias"], int(z["groups"]),
                                           training=(mode == "train"))
    cv.backward(T(z["grad_out"]))
    assert rel_l2(cv.detach().numpy(), z[f"{mode}_cost_volume"]) < 2e-6
    assert rel_l2(np.stack([f.grad.numpy() for f in feats]), z[f"{mode}_grad_features"]) < 2e-5
    assert rel_l2(P["cw"].grad.numpy(), z[f"{mode}_grad_cw"]) < 1e-4
    got_bn = np.array([P["bn_weight"].grad.item(), P["bn_bias"].grad.item()])
    got_fc = np.array([P["fc_weight"].grad.item(), P["fc_bias"].grad.item()])
    assert np.allclose(got_bn, z[f"{mode}_grad_bn"], rtol=2e-4, atol=1e-5)
    assert np.allclose(got_fc, z[f"{mode}_grad_fc"], rtol=2e-4, atol=1e-5)
    if mode == "train":
        rm, rv = float(z["p_bn_mean"]), float(z["p_bn_var"])
        for mean, var in stats:                      # momentum 0.1, unbiased variance, once per source view
            rm, rv = 0.9 * rm + 0.1 * float(mean), 0.9 * rv + 0.1 * float(var)
        assert np.allclose([rm, rv], z["train_running"][:2], rtol=1e-5)
        assert int(z["train_running"][2]) == len(stats)


def test_hypos_by_fit():
    """HyposByFit: the oracle against the reference run in float32 and in float64 (tests/golden/hypos_fit.npz).
    gauss1 is ill conditioned in float32 (the reference's own float32 s is up to 5e-3 away from its float64 run, its
    hypotheses 0.1 mm); the oracle solves the normal equations in extended precision and must sit on the float64 run.
    Everything after the fit (upsampling, range, clamps, hypotheses) is pinned to 1-2 ulp given the reference's s."""
    z = load_golden("hypos_fit")
    s1 = co.hypos_fit(z["prob0"], z["hypos0"], z["depth0"], "gauss1")
    assert (np.abs(s1 - z["s1_f64"]) / np.abs(z["s1_f64"])).max() < 5e-7
    ref_noise = (np.abs(z["s1"] - z["s1_f64"]) / np.abs(z["s1_f64"])).max()
    assert ref_noise > 1e-4                                   # documents why bit parity is not the yardstick here
    h1 = co.hypos_generate(z["depth0"], s1, z["depth_range"], "gauss1", 0.95, 24)
    assert np.abs(h1 - z["hypos1_f64"]).max() < 5e-4 < np.abs(z["hypos1"] - z["hypos1_f64"]).max()
    assert np.abs(co.hypos_generate(z["depth0"], z["s1"], z["depth_range"], "gauss1", 0.95, 24) - z["hypos1"]).max() < 2.5e-4
    s2 = co.hypos_fit(z["prob1"], z["hypos1"], z["depth1"], "laplace")
    assert (np.abs(s2 - z["s2"]) / np.abs(z["s2"])).max() < 2e-6
    h2 = co.hypos_generate(z["depth1"], s2, z["depth_range"], "laplace", 1e-5, 8)
    assert np.abs(h2 - z["hypos2"]).max() < 2.5e-4
    assert h2.shape == z["hypos2"].shape == (2, 8, 48, 64)


def test_hypos_by_fit_gauss0():
    """The curve config.py does not wire (depthhypos.py:127-167), on uniform and on per-pixel hypotheses: the oracle against the
    reference's float64 run (and its float32 run, which is well behaved for this curve), then the hypotheses."""
    z = load_golden("hypos_fit_gauss0")
    s = co.hypos_fit(z["prob0"], z["hypos0"], z["depth0"], "gauss0")
    assert (np.abs(s - z["s_f64"]) / np.abs(z["s_f64"])).max() < 2e-6
    assert (np.abs(s - z["s"]) / np.abs(z["s"])).max() < 5e-6
    h1 = co.hypos_generate(z["depth0"], s, z["depth_range"], "gauss0", 0.95, 24)
    assert h1.shape == z["hypos1"].shape and np.abs(h1 - z["hypos1_f64"]).max() < 5e-4
    sp = co.hypos_fit(z["prob_p"], z["hypos_p"], z["depth_p"], "gauss0")
    assert (np.abs(sp - z["s_p_f64"]) / np.abs(z["s_p_f64"])).max() < 5e-6
    h2 = co.hypos_generate(z["depth_p"], sp, z["depth_range"], "gauss0", 0.9, 8, upsample=False)
    assert np.abs(h2 - z["hypos2_f64"]).max() < 5e-4


@pytest.mark.parametrize("name", ["prob_head_s0", "prob_head_s1", "prob_head_s2"])
def test_prob_conv_and_tail(name):
    """Last layer of the reference's regularisers (input / output of `.prob` captured by hooks) and what CoreNet does
    with it: the conv restatement against the reference's logits, then the chained tail against the reference's."""
    z = load_golden(name)
    logits = co.prob_conv(z["x"], z["weight"])
    logits64 = co.prob_conv(z["x"], z["weight"], prec="f64")
    ref_noise = np.abs(z["logits"] - logits64).max()            # the reference's own float32 distance to float64
    assert np.abs(logits - logits64).max() < 4 * max(ref_noise, 1e-6)
    assert np.abs(logits - z["logits"]).max() < 2e-5            # logits are O(1-10): a few float32 ulps of the sums
    prob = co.softmax_depth(logits)
    assert np.abs(prob - z["prob"]).max() < 2e-5
    depth = co.depth_regression(prob, z["depth_hypos"])
    assert np.abs(depth - z["depth"]).max() < 1e-3 * (935.0 - 425.0) / 47.0
    if "confidence_up" in z.files:
        conf = co.confidence_regress(prob, upsample=2)
        assert_confidence_decisions(conf, z["confidence_up"], z["prob"], name)
    else:
        curve = "gauss1" if name.endswith("s0") else "laplace"
        s = co.hypos_fit(z["prob"], z["depth_hypos"], z["depth"], curve)
        ref = z["s_f64"] if curve == "gauss1" else z["s"]
        assert (np.abs(s - ref) / np.abs(ref)).max() < (5e-3 if curve == "gauss1" else 1e-5)
