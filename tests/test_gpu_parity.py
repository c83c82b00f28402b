"""GPU parity: the CUDA path (through torch.library -> C ABI -> libmdf_b200.so) against
 (1) golden outputs of the unmodified reference (tests/golden/*.npz),
 (2) the CPU oracle on seeded inputs at sizes it finishes in seconds,
 (3) size-independent properties at BASELINE.json's full sizes.

Tolerances (BASELINE.json north_star): cost volume 1e-5 relative (norm-wise: rel-L2 and
max-abs / max|ref|), depth 1e-3 of the stage-0 hypothesis interval, confidence
decisions identical on >= 99.99 % of pixels; index / window work on a given prob volume is bit exact.
"""
import numpy as np
import pytest
import torch

from conftest import assert_confidence_decisions, load_golden, max_abs_over_max, rel_l2
from mdf_net_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

FULL = dict(h0=1152, w0=1600, nviews=5)     # BASELINE.json configs[1]
INTERVAL = (935.0 - 425.0) / 47.0     # stage-0 hypothesis interval (depthhypos.py:33, dtueval.py:47)
COST_TOL = 1e-5
ELEM_TOL = 5e-5


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def params_from(z):
    return {k[2:]: z[k] for k in z.files if k.startswith("p_")}


def run_cost_volume(features, ref_proj, src_projs, hypos, p, G, algo=0):
    from mdf_net_b200 import ops
    f32 = lambda v: cu(np.asarray(v, np.float32).reshape(-1))
    out = ops.cost_volume([cu(f) for f in features], cu(ref_proj), [cu(s) for s in src_projs], cu(hypos),
                          f32(p["cw"]), f32(p["bn_weight"]), f32(p["bn_bias"]), f32(p["bn_mean"]), f32(p["bn_var"]),
                          float(p.get("bn_eps", 1e-5)), f32(p["fc_weight"]), f32(p["fc_bias"]), G, algo)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def assert_cost_close(out, ref, what="", truth=None):
    """rel-L2 < 1e-5 against the float32 reference (north_star's tolerance, norm-wise).  Element-wise, a float32
    evaluation of the reference's own formulae is up to ~1e-4 away from their float64 value on a few elements
    (ulp(800 px) = 6e-5 px per rounding of a coordinate; SURVEY 7.2, tools/diag_parity.py), so the element-wise
    bound is the reference's own noise floor: when `truth` (float64 oracle) is given, the CUDA result must be
    as close to it as the float32 reference is; otherwise a fixed 5e-5."""
    assert out.shape == ref.shape
    assert np.isfinite(out).all(), what
    r, m = rel_l2(out, ref), max_abs_over_max(out, ref)
    assert r < COST_TOL, f"{what}: rel_l2={r:.3g}"
    if truth is None:
        assert m < ELEM_TOL, f"{what}: max_abs/max={m:.3g}"
        return
    floor_l2, floor_max = rel_l2(ref, truth), max_abs_over_max(ref, truth)
    mine_l2, mine_max = rel_l2(out, truth), max_abs_over_max(out, truth)
    assert mine_l2 < 1.5 * floor_l2 + 2e-7, f"{what}: vs float64 rel_l2 {mine_l2:.3g}, reference's own {floor_l2:.3g}"
    assert mine_max < 2.0 * floor_max + 2e-6, f"{what}: vs float64 max {mine_max:.3g}, reference's own {floor_max:.3g}"


# ------------------------------------------------------------------------------------ homo_warping
@pytest.mark.parametrize("name", ["warp_uniform", "warp_pixel"])
def test_warp_golden(name):
    import mdf_net_b200 as mdf
    z = load_golden(name)
    for v in range(z["src_projs"].shape[0]):
        w = mdf.homo_warping(cu(z["src_fea"]), cu(z["src_projs"][v]), cu(z["ref_proj"]), cu(z["depth_hypos"]))
        assert_cost_close(w.cpu().numpy(), z["warped"][v], f"{name}[{v}]")


def test_warp_edge_cases():
    import mdf_net_b200 as mdf
    z = load_golden("warp_edge")
    w = mdf.homo_warping(cu(z["src_fea"]), cu(z["src_proj"]), cu(z["ref_proj"]), cu(z["depth_hypos"])).cpu().numpy()
    ref = z["warped"]
    finite = np.isfinite(z["depth_hypos"].ravel())
    assert rel_l2(w[:, :, finite], ref[:, :, finite]) < COST_TOL
    assert np.all((w[:, :, finite] != 0) == (ref[:, :, finite] != 0))
    # non-finite hypotheses: zeros, as the reference's CUDA grid_sample gives (GridSampler.cuh:140-147)
    assert np.all(w[:, :, ~finite] == 0)
    wi = mdf.homo_warping(cu(z["src_fea"]), cu(z["ref_proj"]), cu(z["ref_proj"]), cu(z["depth_hypos"][:, :1]))
    assert rel_l2(wi.cpu().numpy(), z["warped_identity"]) < COST_TOL


# --------------------------------------------------------------------------------- VectorAggregate
@pytest.mark.parametrize("algo", [0, 2])
@pytest.mark.parametrize("name", ["vecagg_s0", "vecagg_s1", "vecagg_s2", "vecagg_n2"])
def test_vector_aggregate_golden(name, algo):
    z = load_golden(name)
    out = run_cost_volume(z["features"], z["ref_proj"], z["src_projs"], z["depth_hypos"], params_from(z),
                          int(z["groups"]), algo)
    assert_cost_close(out, z["cost_volume"], f"{name} algo={algo}")


def test_vector_aggregate_golden_four_channels_per_group():
    z = load_golden("vecagg_cpg4")       # C/G = 4: only the direct kernel applies
    out = run_cost_volume(z["features"], z["ref_proj"], z["src_projs"], z["depth_hypos"], params_from(z), int(z["groups"]))
    assert_cost_close(out, z["cost_volume"], "cpg4")
    from mdf_net_b200 import _cabi
    with pytest.raises(_cabi.MdfError):   # the staged kernel refuses it instead of falling back silently
        run_cost_volume(z["features"], z["ref_proj"], z["src_projs"], z["depth_hypos"], params_from(z), int(z["groups"]), algo=1)


def test_noise_floor_vs_float64():
    """The CUDA result is as close to a float64 evaluation of the same formulae as the reference is."""
    from oracle import c_oracle as co
    for name in ("vecagg_s0", "vecagg_s1", "vecagg_s2"):
        z = load_golden(name)
        p = params_from(z)
        truth = co.vector_aggregate(list(z["features"]), z["depth_hypos"], p, int(z["groups"]), ref_proj=z["ref_proj"],
                                    src_projs=list(z["src_projs"]), prec="f64")
        out = run_cost_volume(z["features"], z["ref_proj"], z["src_projs"], z["depth_hypos"], p, int(z["groups"]))
        assert rel_l2(out, truth) < 3 * max(rel_l2(z["cost_volume"], truth), 2e-7), name


def stage_case(stage, h0, w0, nviews, batch=1, seed=100):
    H, W = syn.stage_shapes(h0, w0)[stage]
    C, D, G = syn.STAGE_CHANNELS[stage], syn.STAGE_DEPTHS[stage], syn.STAGE_GROUPS[stage]
    K, E = syn.camera_rig(batch, nviews, h0, w0, seed=seed)
    P = syn.projection_matrices(K, E, level_div=2.0 ** (3 - stage))     # scale.py:4-20
    feats = syn.smooth_features(batch, nviews, C, H, W, seed=seed + stage)
    hyp = syn.uniform_hypos(batch, D) if stage == 0 else syn.pixel_hypos(batch, D, H, W, seed=seed + stage)
    return dict(features=feats, ref_proj=P[:, 0], src_projs=[P[:, v] for v in range(1, nviews)], hypos=hyp,
                params=syn.depth_weight_params(G, seed=seed + stage), G=G, C=C, D=D, H=H, W=W)


@pytest.mark.parametrize("stage", [0, 1, 2])
def test_vector_aggregate_vs_oracle_config1(stage):
    """BASELINE.json configs[0] shapes: 640x512, N=3 (the reference's CPU-runnable case)."""
    from oracle import c_oracle as co
    c = stage_case(stage, 512, 640, 3)
    kw = dict(ref_proj=c["ref_proj"], src_projs=c["src_projs"])
    ref = co.vector_aggregate(c["features"], c["hypos"], c["params"], c["G"], **kw)
    truth = co.vector_aggregate(c["features"], c["hypos"], c["params"], c["G"], prec="f64", **kw)
    for algo in (1, 2):
        out = run_cost_volume(c["features"], c["ref_proj"], c["src_projs"], c["hypos"], c["params"], c["G"], algo)
        assert_cost_close(out, ref, f"stage {stage} algo {algo}", truth=truth)


def test_vector_aggregate_batch_and_ragged_sizes():
    """B=2, sizes that are not multiples of the 32-pixel tile, D not a multiple of the plane slab."""
    from oracle import c_oracle as co
    rng = np.random.default_rng(7)
    for (C, G, D, H, W, N) in [(64, 32, 5, 9, 37, 3), (32, 16, 7, 21, 45, 4), (16, 8, 11, 33, 70, 2)]:
        K, E = syn.camera_rig(2, N, H * 8, W * 8, seed=int(rng.integers(1 << 30)))
        P = syn.projection_matrices(K, E, level_div=8.0)
        feats = syn.smooth_features(2, N, C, H, W, seed=5)
        hyp = syn.pixel_hypos(2, D, H, W, seed=6, max_rel_range=0.2)
        p = syn.depth_weight_params(G, seed=8)
        ref = co.vector_aggregate(feats, hyp, p, G, ref_proj=P[:, 0], src_projs=[P[:, v] for v in range(1, N)])
        out = run_cost_volume(feats, P[:, 0], [P[:, v] for v in range(1, N)], hyp, p, G, 1)
        assert_cost_close(out, ref, f"G={G}")


def test_sample_positions_bit_exact():
    """The hot kernel's coordinate chain (shared-reciprocal division, no conversion-pipe floor) gives the
    oracle's sample positions bit for bit -- including behind-the-camera, tiny, huge and non-finite depths."""
    import ctypes
    from mdf_net_b200 import _cabi
    from oracle import c_oracle as co
    lib = _cabi.lib()
    for (h0, w0, stage, seed) in [(512, 640, 1, 5), (1152, 1600, 2, 6), (1152, 1600, 0, 7), (1056, 1920, 2, 8)]:
        c = stage_case(stage, h0, w0, 3, seed=seed)
        H, W, D = c["H"], c["W"], c["D"]
        hyp = c["hypos"][0].reshape(D, -1)
        hyp = hyp[:, 0].copy() if hyp.shape[1] == 1 else c["hypos"][0].copy()
        if hyp.ndim == 1:
            hyp[:6] = [-300.0, 0.0, 1e-3, 1e9, np.inf, np.nan]
        else:
            hyp[0, :4, :6] = np.array([-300.0, 0.0, 1e-3, 1e9, np.inf, np.nan], np.float32)
        for v in range(2):
            rt = co.compose_proj(c["src_projs"][v], c["ref_proj"])[0]
            rx, ry = co.sample_positions(rt, hyp, H, W)
            ix = torch.empty((D, H, W), device="cuda"); iy = torch.empty_like(ix)
            rt_d, hyp_d = cu(rt), cu(hyp)      # keep the device buffers alive across the raw-pointer call
            st = lib.mdf_debug_sample_positions(rt_d.data_ptr(), hyp_d.data_ptr(), int(hyp.ndim == 3), D, H, W,
                                                ix.data_ptr(), iy.data_ptr(), None)
            assert st == 0
            torch.cuda.synchronize()
            for got, ref in ((ix.cpu().numpy(), rx), (iy.cpu().numpy(), ry)):
                same = (got.view(np.uint32) == ref.view(np.uint32)) | (np.isnan(got) & np.isnan(ref))
                assert same.all(), f"{(~same).sum()} of {same.size} positions differ"


def test_rough_hypotheses_need_several_staging_rounds():
    """i.i.d. per-pixel depths (no spatial coherence at all): the samples of one tile scatter over far more
    than one box; the staging loop must still serve every one of them."""
    from oracle import c_oracle as co
    for stage in (1, 2):
        c = stage_case(stage, 256, 320, 4, seed=77)
        c["hypos"] = syn.pixel_hypos(1, c["D"], c["H"], c["W"], seed=78, max_rel_range=0.2, smooth=False)
        ref = co.vector_aggregate(c["features"], c["hypos"], c["params"], c["G"], ref_proj=c["ref_proj"], src_projs=c["src_projs"])
        out = run_cost_volume(c["features"], c["ref_proj"], c["src_projs"], c["hypos"], c["params"], c["G"], 1)
        truth = co.vector_aggregate(c["features"], c["hypos"], c["params"], c["G"], ref_proj=c["ref_proj"],
                                    src_projs=c["src_projs"], prec="f64")
        assert_cost_close(out, ref, f"rough stage {stage}", truth=truth)


def test_eleven_views_like_tanks_and_temples():
    """config.py:119: EvalTanks.nviews = 11 (10 source views); also the 1920x1056 aspect ratio."""
    from oracle import c_oracle as co
    for stage in (0, 2):
        c = stage_case(stage, 1056 // 4, 1920 // 4, 11, seed=900)
        kw = dict(ref_proj=c["ref_proj"], src_projs=c["src_projs"])
        ref = co.vector_aggregate(c["features"], c["hypos"], c["params"], c["G"], **kw)
        truth = co.vector_aggregate(c["features"], c["hypos"], c["params"], c["G"], prec="f64", **kw)
        out = run_cost_volume(c["features"], c["ref_proj"], c["src_projs"], c["hypos"], c["params"], c["G"], 1)
        assert_cost_close(out, ref, f"11 views stage {stage}", truth=truth)


def test_scene_hypotheses_full_size():
    """The benchmark's own stage-1/2 hypotheses (synthetic.scene_hypos) at 1600x1152 N=5 against the oracle."""
    from oracle import c_oracle as co
    c = stage_case(1, FULL["h0"], FULL["w0"], FULL["nviews"], seed=950)
    c["hypos"] = syn.scene_hypos(1, c["D"], c["H"], c["W"], seed=951)
    kw = dict(ref_proj=c["ref_proj"], src_projs=c["src_projs"])
    ref = co.vector_aggregate(c["features"], c["hypos"], c["params"], c["G"], **kw)
    truth = co.vector_aggregate(c["features"], c["hypos"], c["params"], c["G"], prec="f64", **kw)
    out = run_cost_volume(c["features"], c["ref_proj"], c["src_projs"], c["hypos"], c["params"], c["G"])
    assert_cost_close(out, ref, "scene hypotheses stage 1", truth=truth)


def test_corenet_stages_teacher_forced():
    """Replay the three stages of a whole reference CoreNet forward through the drop-in units."""
    import mdf_net_b200 as mdf
    z = load_golden("corenet_64x64_n3")
    mods = torch.nn.ModuleList([mdf.VectorAggregate(g) for g in (32, 16, 8)]).cuda().eval()
    with torch.no_grad():
        for s, m in enumerate(mods):
            bn, fc = z[f"s{s}_bn"], z[f"s{s}_fc"]
            m.depth_weight[0].conv.weight.copy_(cu(z[f"s{s}_cw"]).view(1, -1, 1, 1, 1))
            m.depth_weight[0].bn.weight.fill_(bn[0]); m.depth_weight[0].bn.bias.fill_(bn[1])
            m.depth_weight[0].bn.running_mean.fill_(bn[2]); m.depth_weight[0].bn.running_var.fill_(bn[3])
            m.depth_weight[1].weight.fill_(fc[0]); m.depth_weight[1].bias.fill_(fc[1])
            hyp = cu(z[f"s{s}_depth_hypos"])
            cv = m([cu(f) for f in z[f"s{s}_features"]], cu(z[f"s{s}_ref_proj"]), [cu(p) for p in z[f"s{s}_src_projs"]], hyp)
            assert_cost_close(cv.cpu().numpy(), z[f"s{s}_cost_volume"], f"stage {s}")
            prob, depth = mdf.softmax_regress(cu(z[f"s{s}_logits"]), hyp)
            assert np.abs(prob.cpu().numpy() - z[f"s{s}_prob"]).max() < 3e-7
            assert np.abs(depth.cpu().numpy() - z[f"s{s}_depth"]).max() < 1e-3 * INTERVAL
            d2 = mdf.depth_regression(cu(z[f"s{s}_prob"]), hyp)
            assert np.abs(d2.cpu().numpy() - z[f"s{s}_depth"]).max() < 1e-3 * INTERVAL
        conf = mdf.confidence_regress(cu(z["s2_prob"]))
        conf = torch.nn.functional.interpolate(conf.unsqueeze(1), scale_factor=2, mode="nearest").squeeze(1)   # core.py:76
        assert np.array_equal(conf.cpu().numpy(), z["confidence"])


# ------------------------------------------------------------------------------- variance aggregate
@pytest.mark.parametrize("name", ["varagg_uniform", "varagg_pixel"])
def test_variance_aggregate_golden(name):
    import mdf_net_b200 as mdf
    z = load_golden(name)
    out = mdf.homo_aggregate_by_variance([cu(f) for f in z["features"]], cu(z["ref_proj"]),
                                         [cu(s) for s in z["src_projs"]], cu(z["depth_hypos"]))
    assert_cost_close(out.cpu().numpy(), z["cost_volume"], name)


@pytest.mark.parametrize("C", [12, 16, 32, 64, 72])
def test_variance_aggregate_channel_counts_vs_oracle(C):
    """The channel-quad kernels (C = 16 / 32 / 64: 1 / 2 / 4 lanes per sample, a sample count that does not fill the last
    block), the register-resident kernels (other C <= 64, incl. a count that is not a multiple of 8) and the chunked kernel
    (C > 64) against the oracle."""
    import mdf_net_b200 as mdf
    from oracle import c_oracle as co
    B, N, D, H, W = 2, 4, 5, 20, 28
    K, E = syn.camera_rig(B, N, 8 * H, 8 * W, seed=77)
    P = syn.projection_matrices(K, E, level_div=8.0)
    feats = syn.smooth_features(B, N, C, H, W, seed=78)
    hyp = syn.pixel_hypos(B, D, H, W, seed=79)
    out = mdf.homo_aggregate_by_variance([cu(f) for f in feats], cu(P[:, 0]), [cu(P[:, v]) for v in range(1, N)], cu(hyp))
    ref = co.variance_aggregate(feats, hyp, ref_proj=P[:, 0], src_projs=[P[:, v] for v in range(1, N)])
    assert out.shape == (B, C, D, H, W)
    assert_cost_close(out.cpu().numpy(), ref, f"variance C={C}")


# --------------------------------------------------------------------------------------------- head
@pytest.mark.parametrize("name", ["head_d48", "head_d24", "head_d8"])
def test_head_golden(name):
    import mdf_net_b200 as mdf
    from mdf_net_b200 import ops
    z = load_golden(name)
    for kind in ("uniform", "pixel"):
        hyp = cu(z["hypos_" + kind])
        prob, depth = mdf.softmax_regress(cu(z["logits"]), hyp)
        assert np.abs(prob.cpu().numpy() - z["prob"]).max() < 3e-7
        assert np.abs(depth.cpu().numpy() - z["depth_" + kind]).max() < 1e-3 * INTERVAL
        d = mdf.depth_regression(cu(z["prob"]), hyp)
        assert np.abs(d.cpu().numpy() - z["depth_" + kind]).max() < 1e-3 * INTERVAL
    # window / index work on the reference's own prob volume: bit exact
    assert np.array_equal(mdf.confidence_regress(cu(z["prob"])).cpu().numpy(), z["confidence"])
    assert np.array_equal(ops.confidence(cu(z["prob"]), 4, 1, 2, 2).cpu().numpy(), z["confidence_up"])
    blend = mdf.confidence_regress(cu(z["prob"]), last_confidence=cu(z["last_confidence"]))
    assert np.abs(blend.cpu().numpy() - z["confidence_blend"]).max() < 1e-5
    # fused from logits: same decisions (the probabilities differ from ATen's by <= 1 ulp of exp)
    _, _, conf = mdf.softmax_regress(cu(z["logits"]), cu(z["hypos_pixel"]), want_confidence=True)
    assert_confidence_decisions(conf.cpu().numpy(), z["confidence_up"], z["prob"], name, tol=1e-6)     # tiny fixture: exact count


def test_head_known_answers():
    import mdf_net_b200 as mdf
    z = load_golden("head_known")
    assert np.array_equal(mdf.confidence_regress(cu(z["onehot"])).cpu().numpy(), z["onehot_conf"])
    assert np.array_equal(mdf.confidence_regress(cu(z["uniform"])).cpu().numpy(), z["uniform_conf"])
    # one-hot probability at plane k: depth is exactly hypothesis k
    hyp = syn.uniform_hypos(1, 8)
    d = mdf.depth_regression(cu(z["onehot"]), cu(hyp)).cpu().numpy()
    k = z["onehot"].argmax(1)
    assert np.array_equal(d, hyp.reshape(-1)[k])


def test_head_alternative_window():
    """The n=2 / pad=(0,0,0,0,0,1) variant mentioned at regress.py:9."""
    import mdf_net_b200 as mdf
    from oracle import c_oracle as co
    z = load_golden("head_d8")
    ref = co.confidence_regress(z["prob"], n=2, pad=(0, 0, 0, 0, 0, 1))
    out = mdf.confidence_regress(cu(z["prob"]), n=2, pad=(0, 0, 0, 0, 0, 1))
    assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("stage", [0, 1, 2])
def test_head_vs_oracle_config1(stage):
    import mdf_net_b200 as mdf
    from oracle import c_oracle as co
    H, W = syn.stage_shapes(512, 640)[stage]
    D = syn.STAGE_DEPTHS[stage]
    logits = syn.regulariser_logits(2, D, H, W, seed=50 + stage)
    hyp = syn.uniform_hypos(2, D) if stage == 0 else syn.pixel_hypos(2, D, H, W, seed=60 + stage)
    prob_ref = co.softmax_depth(logits)
    depth_ref = co.depth_regression(prob_ref, hyp)
    conf_ref = co.confidence_regress(prob_ref, upsample=2)
    prob, depth, conf = mdf.softmax_regress(cu(logits), cu(hyp), want_confidence=True)
    assert np.abs(prob.cpu().numpy() - prob_ref).max() < 3e-7
    assert np.abs(depth.cpu().numpy() - depth_ref).max() < 1e-3 * INTERVAL
    same = np.abs(conf.cpu().numpy() - conf_ref) < 1e-6
    assert same.mean() >= 0.9999, f"confidence decisions identical on {same.mean():.6f} of pixels"
    for thr in (0.6, 0.8):      # gipuma prob_threshold / dynamic filter photo mask
        assert ((conf.cpu().numpy() > thr) == (conf_ref > thr)).mean() >= 0.9999
    prob2, depth2 = mdf.softmax_regress(cu(logits), cu(hyp))     # the sliced (warp-shuffle) variant
    assert np.abs(prob2.cpu().numpy() - prob_ref).max() < 3e-7
    assert np.abs(depth2.cpu().numpy() - depth_ref).max() < 1e-3 * INTERVAL


# ------------------------------------------------------------- full-size properties (BASELINE cfg 2)


@pytest.mark.parametrize("stage", [0, 1, 2])
def test_full_size_staged_equals_direct(stage):
    """At 1600x1152 N=5 the oracle is too slow; the TMA-staged kernel and the independent direct kernel
    (both pinned on the oracle at small sizes) must agree."""
    c = stage_case(stage, FULL["h0"], FULL["w0"], FULL["nviews"], seed=200)
    a = run_cost_volume(c["features"], c["ref_proj"], c["src_projs"], c["hypos"], c["params"], c["G"], 1)
    b = run_cost_volume(c["features"], c["ref_proj"], c["src_projs"], c["hypos"], c["params"], c["G"], 2)
    assert rel_l2(a, b) < COST_TOL and max_abs_over_max(a, b) < 2e-4, f"full stage {stage}"
    assert a.min() >= -1e-6 and a.max() <= 1 + 1e-6          # weighted mean of similarities in [0,1]


@pytest.mark.parametrize("stage", [0, 2])
def test_full_size_vs_oracle(stage):
    """The oracle itself at 1600x1152 N=5 (a few seconds per stage with OpenMP)."""
    from oracle import c_oracle as co
    c = stage_case(stage, FULL["h0"], FULL["w0"], FULL["nviews"], seed=300)
    kw = dict(ref_proj=c["ref_proj"], src_projs=c["src_projs"])
    out = run_cost_volume(c["features"], c["ref_proj"], c["src_projs"], c["hypos"], c["params"], c["G"], 1)
    ref = co.vector_aggregate(c["features"], c["hypos"], c["params"], c["G"], **kw)
    truth = co.vector_aggregate(c["features"], c["hypos"], c["params"], c["G"], prec="f64", **kw)
    assert_cost_close(out, ref, f"full stage {stage} vs oracle", truth=truth)


def test_full_size_view_permutation_and_degenerate_features():
    c = stage_case(1, FULL["h0"], FULL["w0"], FULL["nviews"], seed=400)
    a = run_cost_volume(c["features"], c["ref_proj"], c["src_projs"], c["hypos"], c["params"], c["G"])
    perm = [0, 3, 1, 4, 2]
    feats = [c["features"][i] for i in perm]
    projs = [c["src_projs"][i - 1] for i in perm[1:]]
    b = run_cost_volume(feats, c["ref_proj"], projs, c["hypos"], c["params"], c["G"])
    assert rel_l2(a, b) < 1e-6           # weighted mean over views: order only changes rounding
    flat = [np.full_like(f, 0.37) for f in c["features"]]
    h = run_cost_volume(flat, c["ref_proj"], c["src_projs"], c["hypos"], c["params"], c["G"])
    assert np.abs(h - 0.5).max() < 1e-6  # all-equal features: every similarity is 0.5


def test_full_size_head_properties():
    import mdf_net_b200 as mdf
    H, W = syn.stage_shapes(FULL["h0"], FULL["w0"])[2]
    logits = syn.regulariser_logits(1, 8, H, W, seed=70)
    hyp = syn.pixel_hypos(1, 8, H, W, seed=71)
    prob, depth, conf = mdf.softmax_regress(cu(logits), cu(hyp), want_confidence=True)
    prob, depth, conf = prob.cpu().numpy(), depth.cpu().numpy(), conf.cpu().numpy()
    assert np.abs(prob.sum(1) - 1).max() < 1e-6
    assert (depth >= hyp.min(1) - 1e-3).all() and (depth <= hyp.max(1) + 1e-3).all()
    assert conf.shape == (1, 2 * H, 2 * W) and conf.min() >= 0 and conf.max() <= 1 + 1e-6
    assert np.array_equal(conf[:, ::2, ::2], conf[:, 1::2, 1::2])
    # chaining the split API on the fused kernel's own probabilities reproduces the fused result
    c2 = mdf.confidence_regress(cu(prob))
    d2 = mdf.depth_regression(cu(prob), cu(hyp))
    assert (np.abs(c2.cpu().numpy() - conf[:, ::2, ::2]) < 1e-6).mean() >= 0.9999
    assert np.abs(d2.cpu().numpy() - depth).max() < 1e-3 * INTERVAL


def test_tiny_maps_and_maximum_views():
    """Feature maps smaller than one tile / one TMA box, a 1-pixel-wide map (the reference's normalisation divides by
    (W-1)/2 = 0 there: every sample is invalid -> 0.5), and MDF_MAX_VIEWS = 32 views."""
    from oracle import c_oracle as co
    for (H, W, N) in [(2, 3, 3), (5, 33, 2), (3, 1, 3), (6, 8, 32)]:
        C, G, D = 16, 8, 3
        K, E = syn.camera_rig(1, N, H * 8, max(W, 2) * 8, seed=41)
        P = syn.projection_matrices(K, E, level_div=8.0)
        feats = syn.smooth_features(1, N, C, H, W, seed=42)
        hyp = syn.pixel_hypos(1, D, H, W, seed=43)
        p = syn.depth_weight_params(G, seed=44)
        ref = co.vector_aggregate(feats, hyp, p, G, ref_proj=P[:, 0], src_projs=[P[:, v] for v in range(1, N)])
        for algo in (1, 2):
            out = run_cost_volume(feats, P[:, 0], [P[:, v] for v in range(1, N)], hyp, p, G, algo)
            assert np.isfinite(out).all()
            assert np.abs(out - np.nan_to_num(ref, nan=0.5)).max() < 2e-5, (H, W, N, algo)


def test_forward_is_deterministic():
    """No atomics on the data path of the eval forward: two runs give the same bits (also across staging variants
    that change tile / box shapes only)."""
    c = stage_case(2, 512, 640, 4, seed=1234)
    a = run_cost_volume(c["features"], c["ref_proj"], c["src_projs"], c["hypos"], c["params"], c["G"], 1)
    b = run_cost_volume(c["features"], c["ref_proj"], c["src_projs"], c["hypos"], c["params"], c["G"], 1)
    assert np.array_equal(a, b)


# ------------------------------------------------------------------------------------ error handling
def test_errors_and_empty_inputs():
    import mdf_net_b200 as mdf
    from mdf_net_b200 import _cabi, ops
    f = [torch.zeros(1, 16, 8, 8, device="cuda")] * 2
    eye = torch.eye(4, device="cuda")[None]
    with pytest.raises(RuntimeError):
        mdf.homo_warping(f[0], eye, eye, torch.ones(1, 4, 3, 3, device="cuda"))            # hypotheses of another size
    with pytest.raises(RuntimeError):
        mdf.depth_regression(torch.zeros(1, 8, 4, 4, device="cuda"), torch.ones(1, 7, 1, 1, device="cuda"))
    with pytest.raises(RuntimeError):
        mdf.depth_regression(torch.zeros(1, 8, 4, 4, device="cuda", dtype=torch.float64), torch.ones(1, 8, 1, 1, device="cuda"))
    # host pointer through the raw C ABI: rejected, no CPU fallback
    lib = _cabi.lib()
    host = np.zeros(64, np.float32)
    dev = torch.zeros(64, device="cuda")
    st = lib.mdf_depth_regression_fwd(host.ctypes.data, dev.data_ptr(), 0, 1, 4, 4, 4, dev.data_ptr(), None)
    assert st == -5
    # empty batch
    out = mdf.homo_warping(torch.zeros(0, 4, 8, 8, device="cuda"), eye[:0], eye[:0], torch.ones(0, 3, 1, 1, device="cuda"))
    assert out.shape == (0, 4, 3, 8, 8)
    m = mdf.VectorAggregate(8).cuda().eval()
    with torch.no_grad():
        cv = m(f, eye, [eye], torch.full((1, 3, 1, 1), 500.0, device="cuda"))
    assert cv.shape == (1, 8, 3, 8, 8) and torch.allclose(cv, torch.full_like(cv, 0.5))    # zero features
    assert ops.launch_count() > 0


@pytest.mark.parametrize("h0,w0,nviews,stage", [(1056, 1920, 7, 2), (1056, 1920, 7, 0), (1184, 1600, 5, 1)])
def test_other_baseline_shapes_staged_equals_direct(h0, w0, nviews, stage):
    """BASELINE.json configs[3] (Tanks and Temples 1920x1056, N=7) and the crop the shipped DTU loader really uses
    (1600x1184, dtueval.py:34; rows not a multiple of the tile height): staged vs the independent direct kernel."""
    c = stage_case(stage, h0, w0, nviews, seed=400 + stage)
    a = run_cost_volume(c["features"], c["ref_proj"], c["src_projs"], c["hypos"], c["params"], c["G"], 1)
    b = run_cost_volume(c["features"], c["ref_proj"], c["src_projs"], c["hypos"], c["params"], c["G"], 2)
    assert a.shape == (1, c["G"], c["D"], c["H"], c["W"])
    assert rel_l2(a, b) < COST_TOL and max_abs_over_max(a, b) < 2e-4
