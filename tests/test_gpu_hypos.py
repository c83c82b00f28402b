"""GPU parity of the HyposByFit drop-in against the reference's outputs (golden, float32 and float64
runs of the unmodified reference) and the CPU oracle at BASELINE.json's sizes."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from mdf_net_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_hypos_by_fit_golden():
    import mdf_net_b200 as mdf
    z = load_golden("hypos_fit")
    dr = cu(z["depth_range"])
    h0 = mdf.HyposByFit(48, None, 0.0)(None, dr, None, None, upsample=True)
    # stage 0: the reference's own two torch ops; torch's CPU and CUDA elementwise kernels differ by one ulp here
    assert np.abs(h0.cpu().numpy() - z["hypos0"]).max() <= 1.3e-4
    m1 = mdf.HyposByFit(24, "gauss1", 0.95)
    h1 = m1(cu(z["depth0"]), dr, cu(z["prob0"]), cu(z["hypos0"]), upsample=True).cpu().numpy()
    ref_noise = np.abs(z["hypos1"] - z["hypos1_f64"]).max()                            # the reference's own float32 noise: 0.1 mm
    assert np.abs(h1 - z["hypos1_f64"]).max() < 1e-3 < ref_noise                       # we sit on the float64 evaluation
    assert np.abs(h1 - z["hypos1"]).max() < 1.5 * ref_noise
    m2 = mdf.HyposByFit(8, "laplace", 1e-5)
    h2 = m2(cu(z["depth1"]), dr, cu(z["prob1"]), cu(z["hypos1"]), upsample=True).cpu().numpy()
    assert h2.shape == z["hypos2"].shape
    assert np.abs(h2 - z["hypos2"]).max() < 5e-4                                       # well conditioned: ulps of 900 mm
    h2n = m2(cu(z["depth1"]), dr, cu(z["prob1"]), cu(z["hypos1"]), upsample=False)     # the upsample=False branch
    assert h2n.shape == (2, 8, 24, 32)
    with pytest.raises(NotImplementedError):
        mdf.HyposByFit(8, "cauchy", 0.9)(cu(z["depth1"]), dr, cu(z["prob1"]), cu(z["hypos1"]))


def test_hypos_by_fit_gauss0_golden():
    """'gauss0' (depthhypos.py:127-167; config.py does not wire it): the stand-alone fit and generation kernels against the
    reference's float64 and float32 runs, uniform and per-pixel hypotheses, with and without the x2 upsampling."""
    import mdf_net_b200 as mdf
    from mdf_net_b200 import ops
    z = load_golden("hypos_fit_gauss0")
    dr = cu(z["depth_range"])
    s = ops.hypos_fit(cu(z["prob0"]), cu(z["hypos0"]), cu(z["depth0"]), "gauss0").cpu().numpy()
    assert (np.abs(s - z["s_f64"]) / np.abs(z["s_f64"])).max() < 2e-6
    assert (np.abs(s - z["s"]) / np.abs(z["s"])).max() < 5e-6
    h1 = mdf.HyposByFit(24, "gauss0", 0.95)(cu(z["depth0"]), dr, cu(z["prob0"]), cu(z["hypos0"]), upsample=True).cpu().numpy()
    assert h1.shape == z["hypos1"].shape and np.abs(h1 - z["hypos1_f64"]).max() < 1e-3
    sp = ops.hypos_fit(cu(z["prob_p"]), cu(z["hypos_p"]), cu(z["depth_p"]), "gauss0").cpu().numpy()
    assert (np.abs(sp - z["s_p_f64"]) / np.abs(z["s_p_f64"])).max() < 5e-6
    h2 = mdf.HyposByFit(8, "gauss0", 0.9)(cu(z["depth_p"]), dr, cu(z["prob_p"]), cu(z["hypos_p"]), upsample=False).cpu().numpy()
    assert np.abs(h2 - z["hypos2_f64"]).max() < 1e-3
    # the fused tails only take the curves config.py wires
    with pytest.raises(RuntimeError):
        ops.softmax_regress_fit(cu(z["prob0"]).log(), cu(z["hypos0"]), "gauss0")


@pytest.mark.parametrize("stage", [0, 1])
def test_hypos_by_fit_vs_oracle_full_size(stage):
    """Stage 0 -> 1 (gauss1, D=48 uniform hypotheses) and stage 1 -> 2 (laplace, D=24 per-pixel hypotheses) at
    1600x1152: fitted scale and hypotheses against the oracle; properties of the result."""
    import mdf_net_b200 as mdf
    from mdf_net_b200 import ops
    from oracle import c_oracle as co
    H, W = syn.stage_shapes(1152, 1600)[stage]
    D, ND = syn.STAGE_DEPTHS[stage], syn.STAGE_DEPTHS[stage + 1]
    curve, thresh = (("gauss1", 0.95), ("laplace", 1e-5))[stage]
    dr = np.array([[425.0, 935.0]], np.float32)
    logits = syn.regulariser_logits(1, D, H, W, seed=90 + stage, peak=6.0)
    prob = co.softmax_depth(logits)
    hyp = syn.uniform_hypos(1, D) if stage == 0 else syn.scene_hypos(1, D, H, W, seed=91)
    depth = co.depth_regression(prob, hyp)
    s = ops.hypos_fit(cu(prob), cu(hyp), cu(depth), curve).cpu().numpy()
    s_ref = co.hypos_fit(prob, hyp, depth, curve)
    assert (np.abs(s - s_ref) / np.abs(s_ref)).max() < (2e-6 if stage == 0 else 5e-6)
    out = mdf.HyposByFit(ND, curve, thresh)(cu(depth), cu(dr), cu(prob), cu(hyp), upsample=True).cpu().numpy()
    ref = co.hypos_generate(depth, s_ref, dr, curve, thresh, ND)
    assert out.shape == (1, ND, 2 * H, 2 * W)
    assert np.abs(out - ref).max() < 1e-3                                    # mm; 1e-4 of the stage-0 interval
    assert (np.diff(out, axis=1) >= 0).all() and out.min() >= 425.0 and out.max() <= 935.0
    assert (out[:, -1] - out[:, 0]).max() <= 0.2 * (935.0 - 425.0) + 1e-3    # depthhypos.py:59-60
