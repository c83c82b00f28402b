"""GPU parity of the fused regulariser tail: `mdf_prob_head_fwd` (prob conv + softmax + depth
regression + confidence + curve fit, one launch) and `mdf_softmax_regress_fit_fwd` (the same from logits) against the
reference's outputs (goldens made from the unmodified RegularNet_3Scales / RegularNet_4Scales, regress.py and
HyposByFit) and against the CPU oracle at BASELINE.json's sizes."""
import numpy as np
import pytest
import torch

from conftest import assert_confidence_decisions, load_golden
from mdf_net_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

INTERVAL = (935.0 - 425.0) / 47.0


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel(a, b):
    return float((np.abs(a - b) / np.abs(b)).max())


@pytest.mark.parametrize("name", ["prob_head_s0", "prob_head_s1", "prob_head_s2"])
def test_prob_head_golden(name):
    from mdf_net_b200 import ops
    z = load_golden(name)
    last = "confidence_up" in z.files
    curve = "" if last else ("gauss1" if name.endswith("s0") else "laplace")
    logits, prob, depth, conf, s = ops.prob_head(cu(z["x"]), cu(z["weight"]), cu(z["depth_hypos"]), curve, want_logits=True,
                                                 want_prob=True, want_confidence=last)
    assert np.abs(logits.cpu().numpy() - z["logits"]).max() < 2e-5        # a few float32 ulps of sums of O(10) logits
    assert np.abs(prob.cpu().numpy() - z["prob"]).max() < 2e-5
    assert np.abs(depth.cpu().numpy() - z["depth"]).max() < 1e-3 * INTERVAL
    if last:
        c = conf.cpu().numpy()
        assert c.shape == z["confidence_up"].shape
        assert_confidence_decisions(c, z["confidence_up"], z["prob"], name)      # exact flip count, every flip explained
    elif curve == "gauss1":
        ref_noise = rel(z["s"], z["s_f64"])                                   # the reference's own float32 run: 6e-3
        assert rel(s.cpu().numpy(), z["s_f64"]) < 1e-3 < ref_noise
    else:
        assert rel(s.cpu().numpy(), z["s"]) < 1e-4
    # nothing but the requested outputs, and the same numbers, when the logits / probabilities are not materialised
    _, p2, d2, c2, s2 = ops.prob_head(cu(z["x"]), cu(z["weight"]), cu(z["depth_hypos"]), curve, want_logits=False,
                                      want_prob=False, want_confidence=last)
    assert p2.numel() == 0 and torch.equal(d2, depth) and torch.equal(c2, conf) and torch.equal(s2, s)


@pytest.mark.parametrize("stage", [0, 1, 2])
def test_prob_head_vs_oracle_full_size(stage):
    """1600x1152 shapes (c0 = 16 / 8 / 8 feature channels, D = 48 / 24 / 8): the fused launch against the oracle chain
    conv -> softmax -> regression -> confidence / fit."""
    from mdf_net_b200 import ops
    from oracle import c_oracle as co
    H, W = syn.stage_shapes(1152, 1600)[stage]
    D, C = syn.STAGE_DEPTHS[stage], (16, 8, 8)[stage]
    rng = np.random.default_rng(500 + stage)
    x = np.maximum(rng.standard_normal((1, C, D, H, W)), 0).astype(np.float32)      # post-ReLU activations
    w = (rng.standard_normal((1, C, 3, 3, 3)) * 0.35).astype(np.float32)
    hyp = syn.uniform_hypos(1, D) if stage == 0 else syn.scene_hypos(1, D, H, W, seed=91)
    curve = ("gauss1", "laplace", "")[stage]
    logits, prob, depth, conf, s = ops.prob_head(cu(x), cu(w), cu(hyp), curve, want_logits=True, want_prob=True,
                                                 want_confidence=stage == 2)
    co.set_num_threads(co.host_threads())
    lref = co.prob_conv(x, w)
    scale = float(np.abs(lref).max())
    assert np.abs(logits.cpu().numpy() - lref).max() < 4e-6 * max(scale, 1.0)
    pref = co.softmax_depth(lref)
    assert np.abs(prob.cpu().numpy() - pref).max() < 1e-5 * max(scale, 1.0)
    dref = co.depth_regression(pref, hyp)
    assert np.abs(depth.cpu().numpy() - dref).max() < 1e-3 * INTERVAL
    if stage == 2:
        cref = co.confidence_regress(pref, upsample=2)
        c = conf.cpu().numpy()
        assert (np.abs(c - cref) < 1e-4).mean() >= 0.9999
        for thr in (0.6, 0.8):
            assert ((c > thr) == (cref > thr)).mean() >= 0.9999
    else:
        # the fit amplifies the 1e-6 differences of the probabilities: compare on the GPU's own probabilities
        sref = co.hypos_fit(prob.cpu().numpy(), hyp, depth.cpu().numpy(), curve)
        assert rel(s.cpu().numpy(), sref) < 2e-5


@pytest.mark.parametrize("D,curve", [(48, "gauss1"), (24, "laplace"), (8, "laplace"), (12, "gauss1"), (20, "laplace")])
def test_softmax_regress_fit_equals_the_split_path(D, curve):
    """The fit fused into the head (register kernels for D = 8 / 24 / 48, the sweep kernel otherwise) against the
    two-launch path softmax_regress -> hypos_fit and against the oracle."""
    from mdf_net_b200 import ops
    from oracle import c_oracle as co
    B, H, W = 2, 37, 53
    logits = syn.regulariser_logits(B, D, H, W, seed=600 + D, peak=6.0)
    # gauss1 on a nearly flat column divides by c2 -> 0: a different summation / FMA-contraction order of the float64
    # moments shows up at 1e-5 there (the reference's own float32 run is 6e-3 away from its float64 run)
    tol = 1e-4 if curve == "gauss1" else 5e-6
    for hyp in (syn.uniform_hypos(B, D), syn.pixel_hypos(B, D, H, W, seed=601)):
        prob, depth, _ = ops.softmax_regress(cu(logits), cu(hyp), True, False, 4, 1, 2, 2)
        s_split = ops.hypos_fit(prob, cu(hyp), depth, curve).cpu().numpy()
        p2, d2, c2, s = ops.softmax_regress_fit(cu(logits), cu(hyp), curve, want_prob=True)
        assert torch.equal(p2, prob) and torch.equal(d2, depth) and c2.numel() == 0
        assert rel(s.cpu().numpy(), s_split) < tol
        p3, d3, c3, s3 = ops.softmax_regress_fit(cu(logits), cu(hyp), curve, want_prob=False, want_confidence=True)
        assert p3.numel() == 0 and c3.shape == (B, 2 * H, 2 * W)
        assert rel(s3.cpu().numpy(), s_split) < 1e-4     # its own DS = 1 probabilities: ulps away from `prob`
        sref = co.hypos_fit(prob.cpu().numpy(), hyp, depth.cpu().numpy(), curve)
        assert rel(s.cpu().numpy(), sref) < max(tol, 2e-5)


def test_prob_head_ragged_sizes_and_errors():
    from mdf_net_b200 import _cabi, ops
    from oracle import c_oracle as co
    rng = np.random.default_rng(700)
    for (B, C, D, H, W) in [(2, 5, 8, 5, 13), (1, 3, 8, 3, 6), (1, 16, 24, 7, 9), (1, 4, 24, 4, 66), (2, 2, 48, 3, 5), (1, 1, 8, 1, 1)]:
        x = rng.standard_normal((B, C, D, H, W)).astype(np.float32)
        w = (rng.standard_normal((1, C, 3, 3, 3)) * 0.3).astype(np.float32)
        hyp = syn.pixel_hypos(B, D, H, W, seed=701) if H * W > 1 else syn.uniform_hypos(B, D).reshape(B, D, 1, 1)
        logits, prob, depth, conf, s = ops.prob_head(cu(x), cu(w), cu(hyp), "laplace", want_logits=True, want_prob=True,
                                                     want_confidence=True)
        lref = co.prob_conv(x, w)
        assert np.abs(logits.cpu().numpy() - lref).max() < 1e-5, (B, C, D, H, W)
        pref = co.softmax_depth(lref)
        assert np.abs(prob.cpu().numpy() - pref).max() < 1e-5
        assert np.abs(depth.cpu().numpy() - co.depth_regression(pref, hyp)).max() < 1e-3 * INTERVAL
        assert conf.shape == (B, 2 * H, 2 * W)
    # a view (misaligned planes) goes through the scalar variant
    x = cu(rng.standard_normal((1, 4, 8, 6, 9)).astype(np.float32))[:, :, :, :, 1:]
    w = cu((rng.standard_normal((1, 4, 3, 3, 3)) * 0.3).astype(np.float32))
    hyp = cu(syn.pixel_hypos(1, 8, 6, 8, seed=702))
    lg = ops.prob_head(x, w, hyp, "", want_logits=True)[0]
    assert np.abs(lg.cpu().numpy() - co.prob_conv(x.cpu().numpy(), w.cpu().numpy())).max() < 1e-5
    # unsupported depth, empty input
    with pytest.raises(_cabi.MdfError):
        ops.prob_head(cu(np.zeros((1, 4, 10, 4, 4), np.float32)), cu(np.zeros((1, 4, 3, 3, 3), np.float32)),
                      cu(syn.uniform_hypos(1, 10)), "")
    out = ops.prob_head(cu(np.zeros((0, 4, 8, 4, 4), np.float32)), cu(np.zeros((1, 4, 3, 3, 3), np.float32)),
                        cu(np.zeros((0, 8, 1, 1), np.float32)), "")
    assert out[2].shape == (0, 4, 4)
    with pytest.raises(RuntimeError):
        ops.prob_head(torch.zeros((1, 4, 8, 4, 4)), torch.zeros((1, 4, 3, 3, 3)), torch.zeros((1, 8, 1, 1)), "")
