"""The whole stage loop of CoreNet.forward (net/core.py:45-77) with every hot-path unit replaced by this package --
HyposByFit -> VectorAggregate -> [stand-in regulariser] -> softmax / depth regression -> ... -> confidence -- against
the same chain evaluated with the CPU oracle.  The stand-in regulariser (the 3-D CNN is out of scope) is a fixed,
smooth function of the cost volume, applied identically on both sides."""
import numpy as np
import pytest
import torch

from conftest import assert_confidence_decisions
from mdf_net_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

INTERVAL = (935.0 - 425.0) / 47.0
CURVES = [None, "gauss1", "laplace"]          # config.py:200
THRESH = (0.0, 0.95, 1e-5)                    # config.py:201


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


GAIN = 20.0     # mean similarity - 0.5 lies in +-0.3: logits of +-6, the range of a trained regulariser's output


def stand_in_regulariser(cost_volume, xp):
    """logits (B,D,H,W): sharp where the mean similarity over the groups is high."""
    return (cost_volume.mean(1) - 0.5) * GAIN if xp is torch else ((cost_volume.mean(1) - 0.5) * np.float32(GAIN)).astype(np.float32)


def test_stage_loop_matches_the_oracle_chain():
    import mdf_net_b200 as mdf
    from oracle import c_oracle as co
    h0, w0, N, B = 256, 320, 4, 1
    K, E = syn.camera_rig(B, N, h0, w0, seed=77)
    dr = np.array([[425.0, 935.0]], np.float32)
    feats = [syn.smooth_features(B, N, syn.STAGE_CHANNELS[s], *syn.stage_shapes(h0, w0)[s], seed=80 + s) for s in range(3)]
    params = [syn.depth_weight_params(syn.STAGE_GROUPS[s], seed=90 + s) for s in range(3)]

    # ---- this package, on the GPU (the loop of core.py:45-65) ----
    mods = []
    for s in range(3):
        m = mdf.VectorAggregate(syn.STAGE_GROUPS[s]).cuda().eval()
        p, dw = params[s], m.depth_weight
        with torch.no_grad():
            dw[0].conv.weight.copy_(cu(p["cw"]).view(1, -1, 1, 1, 1))
            dw[0].bn.weight.fill_(float(p["bn_weight"])); dw[0].bn.bias.fill_(float(p["bn_bias"]))
            dw[0].bn.running_mean.fill_(float(p["bn_mean"])); dw[0].bn.running_var.fill_(float(p["bn_var"]))
            dw[1].weight.fill_(float(p["fc_weight"])); dw[1].bias.fill_(float(p["fc_bias"]))
        mods.append(m)
    hypos_units = [mdf.HyposByFit(syn.STAGE_DEPTHS[s], CURVES[s], THRESH[s]) for s in range(3)]
    depth = hyp = prob = None
    gpu = {}
    with torch.no_grad():
        for s in range(3):
            P = syn.projection_matrices(K, E, 2.0 ** (3 - s))
            hyp = hypos_units[s](depth, cu(dr), prob, hyp, upsample=True)
            cv = mods[s]([cu(f) for f in feats[s]], cu(P[:, 0]), [cu(P[:, v]) for v in range(1, N)], hyp)
            prob, depth = mdf.softmax_regress(stand_in_regulariser(cv, torch), hyp)
            gpu[s] = (hyp.cpu().numpy(), depth.cpu().numpy())
        conf = mdf.confidence_regress(prob)
        conf = torch.nn.functional.interpolate(conf.unsqueeze(1), scale_factor=2, mode="nearest").squeeze(1).cpu().numpy()

    # ---- the oracle, on the CPU ----
    depth = hyp = prob = None
    for s in range(3):
        P = syn.projection_matrices(K, E, 2.0 ** (3 - s))
        if s == 0:
            hyp = syn.uniform_hypos(B, syn.STAGE_DEPTHS[0])
        else:
            sc = co.hypos_fit(prob, hyp, depth, CURVES[s])
            hyp = co.hypos_generate(depth, sc, dr, CURVES[s], THRESH[s], syn.STAGE_DEPTHS[s])
        cv = co.vector_aggregate(feats[s], hyp, params[s], syn.STAGE_GROUPS[s], ref_proj=P[:, 0], src_projs=[P[:, v] for v in range(1, N)])
        prob = co.softmax_depth(stand_in_regulariser(cv, np))
        depth = co.depth_regression(prob, hyp)
        g_hyp, g_depth = gpu[s]
        # hypotheses and depth of every stage: 1e-3 of the stage-0 interval, chained errors included
        assert np.abs(g_hyp.reshape(hyp.shape) - hyp).max() < 1e-3 * INTERVAL * (1 if s < 2 else 2), s
        assert np.abs(g_depth - depth).max() < 1e-3 * INTERVAL * (1 if s < 2 else 2), s
    conf_ref = co.confidence_regress(prob, upsample=2)
    # confidence: an exact count of the differing pixels / mask decisions at 0.6 and 0.8, each one explained as a
    # truncation flip of the expected index under the chained float32 noise (conftest.assert_confidence_decisions)
    assert_confidence_decisions(conf, conf_ref, prob, "stage loop vs oracle chain", index_noise=1e-3)
