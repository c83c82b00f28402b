"""`torch.library` custom ops (`mdfnet_b200::*`) over the C ABI of libmdf_b200.so.

Each op is a thin shim: it validates dtypes / devices, allocates the output and the scratch
workspace with torch (the C library allocates nothing), and passes raw device pointers, sizes and
torch's *current* CUDA stream across the ABI.  Ops are registered for CUDA only: a CPU tensor is a
hard error, there is no fallback (BASELINE.json north_star).

Reference call sites replaced (file:line into the reference checkout):
  cost_volume       net/unit/homoaggregate.py:25-46   VectorAggregate.forward (eval-mode BN)
  fpn_out_prepped   net/unit/backbone.py:43-45,59-63  the FPN's 1x1 output convolutions, emitting the hot kernel's layout
  cost_volume_prepped  the same cost volume from those maps (optional fast entry, no layout pass)
  homo_warp         net/unit/base.py:85-126           homo_warping
  variance_volume   net/unit/homoaggregate.py:49-69   homo_aggregate_by_variance
  softmax_regress   net/unit/regular.py:67-69,130-133 + net/unit/regress.py:5-25
  prob_head         net/unit/regular.py:43,67-69 / :110,130-133 (prob conv + softmax) + regress + curve fit, one launch
  softmax_regress_fit  the same + net/unit/depthhypos.py:78-125,169-215 (the curve fit of the next stage's HyposByFit)
  depth_regression  net/unit/regress.py:5-7
  confidence        net/unit/regress.py:9-25 (+ core.py:75-77 nearest upsample)
  geo_filter        tools/filter/dynamic_filter_gpu.py:57-100,161-237 (geometric-consistency filter, post-processing)
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
from torch import Tensor

from . import _cabi

__all__ = ["cost_volume", "cost_volume_prepped", "fpn_out_prepped", "homo_warp", "variance_volume", "softmax_regress", "softmax_regress_fit", "prob_head", "depth_regression",
           "confidence", "hypos_fit", "hypos_generate", "geo_filter",
           "launch_count", "reset_launch_count", "time_next_hot_kernel"]

# kernels launched through this module since the last reset (bench.py's `gpu_launches`)
_launches = 0

# kernel launches per C-ABI call (staged: setup + prep + hot kernel; direct: setup + kernel; see csrc/*.cu)
_LAUNCHES_STAGED, _LAUNCHES_DIRECT, _LAUNCHES_SIMPLE = 3, 2, 1


def launch_count() -> int:
    return _launches


def reset_launch_count() -> None:
    global _launches
    _launches = 0


def _count(n: int) -> None:
    global _launches
    _launches += n


# (start, stop) torch.cuda.Event pair the NEXT cost_volume call of this process records around its hot kernel alone
# (mdf_cost_volume_fwd_ex's timing hook; bench.py's live roofline numbers).  Python-side state: the library keeps none.
_hot_events = None


def time_next_hot_kernel(start: "torch.cuda.Event", stop: "torch.cuda.Event") -> None:
    global _hot_events
    start.record(); stop.record()        # materialise the handles
    _hot_events = (start, stop)


def _stream(t: Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream


def _f32c(t: Tensor, what: str) -> Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"mdfnet_b200: {what} must be a CUDA tensor (there is no CPU fallback), got {t.device}")
    if t.dtype != torch.float32:
        raise RuntimeError(f"mdfnet_b200: {what} must be float32, got {t.dtype}")
    return t.contiguous()


def _hypos(h: Tensor, B: int, H: int, W: int) -> Tuple[Tensor, int, int]:
    """(B,D,1,1) -> per_pixel 0; (B,D,H,W) -> per_pixel 1 (base.py:94 reads D,H,W from the hypotheses)."""
    if h.dim() != 4 or h.shape[0] != B:
        raise RuntimeError(f"mdfnet_b200: depth_hypos must be (B,D,1,1) or (B,D,H,W), got {tuple(h.shape)}")
    D = h.shape[1]
    if h.shape[2] == 1 and h.shape[3] == 1 and not (H == 1 and W == 1):
        return _f32c(h, "depth_hypos"), D, 0
    if h.shape[2] == H and h.shape[3] == W:
        return _f32c(h, "depth_hypos"), D, 1
    raise RuntimeError(f"mdfnet_b200: depth_hypos {tuple(h.shape)} does not match the feature size {(H, W)}")


def _workspace(nbytes: int, device) -> Tensor:
    # torch's caching allocator returns >= 512-byte aligned blocks; stream-ordered reuse is safe
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _check_views(features: Sequence[Tensor], src_projs: Sequence[Tensor]):
    if len(features) < 2 or len(src_projs) != len(features) - 1:
        raise RuntimeError(f"mdfnet_b200: need N >= 2 feature maps and N-1 source projections, got "
                           f"{len(features)} and {len(src_projs)}")
    shape = features[0].shape
    if features[0].dim() != 4 or any(f.shape != shape for f in features):
        raise RuntimeError("mdfnet_b200: all feature maps must share one (B,C,H,W) shape")


# ------------------------------------------------------------------------------------- cost volume
@torch.library.custom_op("mdfnet_b200::cost_volume", mutates_args=(), device_types="cuda")
def cost_volume(features: List[Tensor], ref_proj: Tensor, src_projs: List[Tensor], depth_hypos: Tensor,
                conv_weight: Tensor, bn_weight: Tensor, bn_bias: Tensor, bn_mean: Tensor, bn_var: Tensor,
                bn_eps: float, fc_weight: Tensor, fc_bias: Tensor, groups: int, algo: int = 0) -> Tensor:
    """Fused plane-sweep cost volume (B,G,D,H,W); eval-mode depth_weight.  algo: 0 auto, 1 staged, 2 direct."""
    _check_views(features, src_projs)
    feats = [_f32c(f, "features") for f in features]
    B, C, H, W = feats[0].shape
    hyp, D, per_pixel = _hypos(depth_hypos, B, H, W)
    projs = [_f32c(p, "src_projs") for p in src_projs]
    refp = _f32c(ref_proj, "ref_proj")
    params = [_f32c(t, n) for t, n in ((conv_weight, "conv_weight"), (bn_weight, "bn_weight"), (bn_bias, "bn_bias"),
                                        (bn_mean, "bn_mean"), (bn_var, "bn_var"), (fc_weight, "fc_weight"),
                                        (fc_bias, "fc_bias"))]
    if params[0].numel() != groups:
        raise RuntimeError(f"mdfnet_b200: conv_weight has {params[0].numel()} elements, expected groups={groups}")
    lib = _cabi.lib()
    N = len(feats)
    out = torch.empty((B, groups, D, H, W), dtype=torch.float32, device=feats[0].device)
    ws_bytes = lib.mdf_cost_volume_workspace_bytes(B, N, C, groups, D, H, W)
    ws = _workspace(ws_bytes, out.device)
    global _hot_events
    ev, _hot_events = _hot_events, None
    st = lib.mdf_cost_volume_fwd_ex(
        _cabi.ptr_array([f.data_ptr() for f in feats]), N, refp.data_ptr(),
        _cabi.ptr_array([p.data_ptr() for p in projs]), hyp.data_ptr(), per_pixel,
        params[0].data_ptr(), params[1].data_ptr(), params[2].data_ptr(), params[3].data_ptr(), params[4].data_ptr(),
        float(bn_eps), params[5].data_ptr(), params[6].data_ptr(),
        B, C, groups, D, H, W, out.data_ptr(), ws.data_ptr(), ws.numel(), int(algo),
        ev[0].cuda_event if ev else None, ev[1].cuda_event if ev else None, _stream(out))
    _cabi.check("mdf_cost_volume_fwd", st)
    staged = algo != 2 and C == 2 * groups and groups in (8, 16, 32)
    _count(_LAUNCHES_STAGED if staged else _LAUNCHES_DIRECT)
    return out


@cost_volume.register_fake
def _(features, ref_proj, src_projs, depth_hypos, conv_weight, bn_weight, bn_bias, bn_mean, bn_var, bn_eps,
      fc_weight, fc_bias, groups, algo=0):
    B, _, H, W = features[0].shape
    return features[0].new_empty((B, groups, depth_hypos.shape[1], H, W))


# ------------------------------------------------------------- FPN hand-off (optional fast entry)
@torch.library.custom_op("mdfnet_b200::fpn_out_prepped", mutates_args=(), device_types="cuda")
def fpn_out_prepped(x: Tensor, out_weight: Tensor, groups: int, depth_weight_conv: Tensor, is_reference: bool) -> Tuple[Tensor, Tensor]:
    """The FPN's bias-free 1x1 output convolution (backbone.py:43-45, 59-63) of ONE view, emitting the hot kernel's input
    layout instead of NCHW features.  x (B,Cin,H,W), out_weight (2G,Cin[,1,1]).  Source view: returns (S4, empty) with
    S4 (B,G/4,H,W,4) = (y[2g+1]-y[2g])*log2(e); reference view: returns (Q4, CQ4), Q4 = 2*sigmoid(y[2g]-y[2g+1])-1 and
    CQ4 = depth_weight.0.conv.weight[g] * Q4."""
    xx = _f32c(x, "x")
    B, Cin, H, W = xx.shape
    w = _f32c(out_weight, "out_weight").reshape(out_weight.shape[0], -1)
    if w.shape != (2 * groups, Cin):
        raise RuntimeError(f"mdfnet_b200: out_weight must be (2*groups, Cin) = {(2 * groups, Cin)}, got {tuple(w.shape)}")
    J = groups // 4
    a = torch.empty((B, J, H, W, 4), dtype=torch.float32, device=xx.device)
    b = torch.empty((B, J, H, W, 4), dtype=torch.float32, device=xx.device) if is_reference else torch.empty(0, device=xx.device)
    cw = _f32c(depth_weight_conv, "depth_weight_conv").reshape(-1) if is_reference else None
    st = _cabi.lib().mdf_fpn_out_prepped_fwd(xx.data_ptr(), w.data_ptr(), B, Cin, groups, H, W,
                                             cw.data_ptr() if is_reference else None,
                                             None if is_reference else a.data_ptr(),
                                             a.data_ptr() if is_reference else None, b.data_ptr() if is_reference else None, _stream(xx))
    _cabi.check("mdf_fpn_out_prepped_fwd", st)
    _count(_LAUNCHES_SIMPLE)
    return a, b


@fpn_out_prepped.register_fake
def _(x, out_weight, groups, depth_weight_conv, is_reference):
    B, _, H, W = x.shape
    a = x.new_empty((B, groups // 4, H, W, 4))
    return a, (x.new_empty((B, groups // 4, H, W, 4)) if is_reference else x.new_empty(0))


@torch.library.custom_op("mdfnet_b200::cost_volume_prepped", mutates_args=(), device_types="cuda")
def cost_volume_prepped(s4: Tensor, q4: Tensor, cq4: Tensor, ref_proj: Tensor, src_projs: List[Tensor], depth_hypos: Tensor,
                        conv_weight: Tensor, bn_weight: Tensor, bn_bias: Tensor, bn_mean: Tensor, bn_var: Tensor,
                        bn_eps: float, fc_weight: Tensor, fc_bias: Tensor, groups: int) -> Tensor:
    """The fused cost volume from prepared maps: s4 (V,B,G/4,H,W,4) of the V source views, q4 / cq4 (B,G/4,H,W,4) of the
    reference view (ops.fpn_out_prepped).  Same result as ops.cost_volume on the NCHW features, no layout pass."""
    S, Q, CQ = _f32c(s4, "s4"), _f32c(q4, "q4"), _f32c(cq4, "cq4")
    if S.dim() != 6 or Q.dim() != 5 or S.shape[1:] != Q.shape or Q.shape != CQ.shape or S.shape[2] * 4 != groups or S.shape[-1] != 4:
        raise RuntimeError(f"mdfnet_b200: s4 {tuple(S.shape)} / q4 {tuple(Q.shape)} / cq4 {tuple(CQ.shape)} do not fit groups={groups}")
    V, B, _, H, W, _ = S.shape
    if len(src_projs) != V:
        raise RuntimeError(f"mdfnet_b200: {V} source views in s4, {len(src_projs)} source projections")
    hyp, D, per_pixel = _hypos(depth_hypos, B, H, W)
    projs = [_f32c(p, "src_projs") for p in src_projs]
    refp = _f32c(ref_proj, "ref_proj")
    params = [_f32c(t, n) for t, n in ((conv_weight, "conv_weight"), (bn_weight, "bn_weight"), (bn_bias, "bn_bias"),
                                        (bn_mean, "bn_mean"), (bn_var, "bn_var"), (fc_weight, "fc_weight"), (fc_bias, "fc_bias"))]
    lib = _cabi.lib()
    out = torch.empty((B, groups, D, H, W), dtype=torch.float32, device=S.device)
    ws = _workspace(lib.mdf_cost_volume_prepped_workspace_bytes(B, V + 1), out.device)
    st = lib.mdf_cost_volume_fwd_prepped(
        S.data_ptr(), Q.data_ptr(), CQ.data_ptr(), V + 1, refp.data_ptr(), _cabi.ptr_array([p.data_ptr() for p in projs]),
        hyp.data_ptr(), per_pixel, params[0].data_ptr(), params[1].data_ptr(), params[2].data_ptr(), params[3].data_ptr(),
        params[4].data_ptr(), float(bn_eps), params[5].data_ptr(), params[6].data_ptr(), B, groups, D, H, W,
        out.data_ptr(), ws.data_ptr(), ws.numel(), _stream(out))
    _cabi.check("mdf_cost_volume_fwd_prepped", st)
    _count(_LAUNCHES_STAGED - 1)              # setup + hot kernel: no layout pass
    return out


@cost_volume_prepped.register_fake
def _(s4, q4, cq4, ref_proj, src_projs, depth_hypos, conv_weight, bn_weight, bn_bias, bn_mean, bn_var, bn_eps, fc_weight, fc_bias, groups):
    return s4.new_empty((s4.shape[1], groups, depth_hypos.shape[1], s4.shape[3], s4.shape[4]))


# -------------------------------------------------------------------------------------- homo_warp
@torch.library.custom_op("mdfnet_b200::homo_warp", mutates_args=(), device_types="cuda")
def homo_warp(src_fea: Tensor, src_proj: Tensor, ref_proj: Tensor, depth_hypos: Tensor) -> Tensor:
    """homo_warping (base.py:85-126): (B,C,H,W) -> (B,C,D,H,W)."""
    f = _f32c(src_fea, "src_fea")
    if f.dim() != 4:
        raise RuntimeError("mdfnet_b200: src_fea must be (B,C,H,W)")
    B, C, H, W = f.shape
    hyp, D, per_pixel = _hypos(depth_hypos, B, H, W)
    sp, rp = _f32c(src_proj, "src_proj"), _f32c(ref_proj, "ref_proj")
    lib = _cabi.lib()
    out = torch.empty((B, C, D, H, W), dtype=torch.float32, device=f.device)
    ws = _workspace(lib.mdf_homo_warp_workspace_bytes(B), f.device)
    st = lib.mdf_homo_warp_fwd(f.data_ptr(), sp.data_ptr(), rp.data_ptr(), hyp.data_ptr(), per_pixel, B, C, D, H, W,
                               out.data_ptr(), ws.data_ptr(), ws.numel(), _stream(out))
    _cabi.check("mdf_homo_warp_fwd", st)
    _count(_LAUNCHES_DIRECT)
    return out


@homo_warp.register_fake
def _(src_fea, src_proj, ref_proj, depth_hypos):
    B, C, H, W = src_fea.shape
    return src_fea.new_empty((B, C, depth_hypos.shape[1], H, W))


# -------------------------------------------------------------------------------- variance volume
@torch.library.custom_op("mdfnet_b200::variance_volume", mutates_args=(), device_types="cuda")
def variance_volume(features: List[Tensor], ref_proj: Tensor, src_projs: List[Tensor], depth_hypos: Tensor) -> Tensor:
    """homo_aggregate_by_variance (homoaggregate.py:49-69): (B,C,D,H,W)."""
    _check_views(features, src_projs)
    feats = [_f32c(f, "features") for f in features]
    B, C, H, W = feats[0].shape
    hyp, D, per_pixel = _hypos(depth_hypos, B, H, W)
    projs = [_f32c(p, "src_projs") for p in src_projs]
    refp = _f32c(ref_proj, "ref_proj")
    lib = _cabi.lib()
    N = len(feats)
    out = torch.empty((B, C, D, H, W), dtype=torch.float32, device=feats[0].device)
    ws = _workspace(lib.mdf_variance_volume_workspace_bytes(B, N, C, D, H, W), out.device)
    st = lib.mdf_variance_volume_fwd(_cabi.ptr_array([f.data_ptr() for f in feats]), N, refp.data_ptr(),
                                     _cabi.ptr_array([p.data_ptr() for p in projs]), hyp.data_ptr(), per_pixel,
                                     B, C, D, H, W, out.data_ptr(), ws.data_ptr(), ws.numel(), _stream(out))
    _cabi.check("mdf_variance_volume_fwd", st)
    _count(_LAUNCHES_DIRECT)
    return out


@variance_volume.register_fake
def _(features, ref_proj, src_projs, depth_hypos):
    B, C, H, W = features[0].shape
    return features[0].new_empty((B, C, depth_hypos.shape[1], H, W))


# ------------------------------------------------------------------------------------------- head
_CURVES = {"gauss1": 1, "laplace": 2}      # depthhypos.py:44-47 (config.py:200 wires None, "gauss1", "laplace"): the fused tails
_FIT_CURVES = {**_CURVES, "gauss0": 3}     # the stand-alone fit / generation also cover "gauss0" (depthhypos.py:42-43, 127-167)

def _prob(t: Tensor, what: str) -> Tensor:
    t = _f32c(t, what)
    if t.dim() != 4:
        raise RuntimeError(f"mdfnet_b200: {what} must be (B,D,H,W), got {tuple(t.shape)}")
    return t


def _head_hypos(h: Tensor, B: int, D: int, H: int, W: int) -> Tuple[Tensor, int]:
    hyp, Dh, per_pixel = _hypos(h, B, H, W)
    if Dh != D:
        raise RuntimeError(f"mdfnet_b200: depth_hypos has {Dh} planes, the volume has {D}")
    return hyp, per_pixel


@torch.library.custom_op("mdfnet_b200::softmax_regress", mutates_args=(), device_types="cuda")
def softmax_regress(logits: Tensor, depth_hypos: Tensor, want_prob: bool = True, want_confidence: bool = False,
                    n: int = 4, pad_front: int = 1, pad_back: int = 2, upsample: int = 2) -> Tuple[Tensor, Tensor, Tensor]:
    """softmax over D + depth expectation (+ confidence) in one pass.  Returns (prob, depth, confidence);
    outputs that were not requested are empty tensors."""
    x = _prob(logits, "logits")
    B, D, H, W = x.shape
    hyp, per_pixel = _head_hypos(depth_hypos, B, D, H, W)
    dev = x.device
    prob = torch.empty_like(x) if want_prob else torch.empty(0, device=dev)
    depth = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    conf = torch.empty((B, H * upsample, W * upsample), dtype=torch.float32, device=dev) if want_confidence \
        else torch.empty(0, device=dev)
    st = _cabi.lib().mdf_softmax_regress_fwd(
        x.data_ptr(), hyp.data_ptr(), per_pixel, B, D, H, W,
        prob.data_ptr() if want_prob else None, depth.data_ptr(), conf.data_ptr() if want_confidence else None,
        n, pad_front, pad_back, upsample, _stream(x))
    _cabi.check("mdf_softmax_regress_fwd", st)
    _count(_LAUNCHES_SIMPLE)
    return prob, depth, conf


@softmax_regress.register_fake
def _(logits, depth_hypos, want_prob=True, want_confidence=False, n=4, pad_front=1, pad_back=2, upsample=2):
    B, D, H, W = logits.shape
    return (torch.empty_like(logits) if want_prob else logits.new_empty(0), logits.new_empty((B, H, W)),
            logits.new_empty((B, H * upsample, W * upsample)) if want_confidence else logits.new_empty(0))


@torch.library.custom_op("mdfnet_b200::softmax_regress_fit", mutates_args=(), device_types="cuda")
def softmax_regress_fit(logits: Tensor, depth_hypos: Tensor, curve: str, want_prob: bool = False,
                        want_confidence: bool = False, n: int = 4, pad_front: int = 1, pad_back: int = 2,
                        upsample: int = 2) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """softmax over D + depth expectation (+ confidence) + the per-pixel curve fit of HyposByFit, one pass over the
    logits.  Returns (prob, depth, confidence, s); outputs that were not requested are empty tensors."""
    if curve not in _CURVES:
        raise RuntimeError(f"mdfnet_b200: curve must be one of {sorted(_CURVES)}, got {curve!r}")
    x = _prob(logits, "logits")
    B, D, H, W = x.shape
    hyp, per_pixel = _head_hypos(depth_hypos, B, D, H, W)
    dev = x.device
    prob = torch.empty_like(x) if want_prob else torch.empty(0, device=dev)
    depth = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    s = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    conf = torch.empty((B, H * upsample, W * upsample), dtype=torch.float32, device=dev) if want_confidence \
        else torch.empty(0, device=dev)
    st = _cabi.lib().mdf_softmax_regress_fit_fwd(
        x.data_ptr(), hyp.data_ptr(), per_pixel, B, D, H, W,
        prob.data_ptr() if want_prob else None, depth.data_ptr(), conf.data_ptr() if want_confidence else None,
        n, pad_front, pad_back, upsample, _CURVES[curve], s.data_ptr(), _stream(x))
    _cabi.check("mdf_softmax_regress_fit_fwd", st)
    _count(_LAUNCHES_SIMPLE)
    return prob, depth, conf, s


@softmax_regress_fit.register_fake
def _(logits, depth_hypos, curve, want_prob=False, want_confidence=False, n=4, pad_front=1, pad_back=2, upsample=2):
    B, D, H, W = logits.shape
    return (torch.empty_like(logits) if want_prob else logits.new_empty(0), logits.new_empty((B, H, W)),
            logits.new_empty((B, H * upsample, W * upsample)) if want_confidence else logits.new_empty(0),
            logits.new_empty((B, H, W)))


@torch.library.custom_op("mdfnet_b200::prob_head", mutates_args=(), device_types="cuda")
def prob_head(x: Tensor, prob_weight: Tensor, depth_hypos: Tensor, curve: str = "", want_logits: bool = False,
              want_prob: bool = True, want_confidence: bool = False, n: int = 4, pad_front: int = 1, pad_back: int = 2,
              upsample: int = 2, algo: int = 0) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Conv3d(c0,1,3,pad=1,no bias) of the regulariser's last feature volume x (B,c0,D,H,W) + softmax over D + depth
    expectation (+ confidence) (+ curve fit, curve in {"gauss1","laplace"}), one launch.  Returns
    (logits, prob, depth, confidence, s); outputs that were not requested are empty tensors."""
    if curve and curve not in _CURVES:
        raise RuntimeError(f"mdfnet_b200: curve must be '' or one of {sorted(_CURVES)}, got {curve!r}")
    v = _f32c(x, "x")
    if v.dim() != 5:
        raise RuntimeError(f"mdfnet_b200: x must be (B,C,D,H,W), got {tuple(v.shape)}")
    B, C, D, H, W = v.shape
    w = _f32c(prob_weight, "prob_weight")
    if w.numel() != C * 27:
        raise RuntimeError(f"mdfnet_b200: prob_weight must be (1,{C},3,3,3), got {tuple(w.shape)}")
    hyp, per_pixel = _head_hypos(depth_hypos, B, D, H, W)
    dev = v.device
    empty = lambda: torch.empty(0, device=dev)
    logits = torch.empty((B, D, H, W), dtype=torch.float32, device=dev) if want_logits else empty()
    prob = torch.empty((B, D, H, W), dtype=torch.float32, device=dev) if want_prob else empty()
    depth = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    conf = torch.empty((B, H * upsample, W * upsample), dtype=torch.float32, device=dev) if want_confidence else empty()
    s = torch.empty((B, H, W), dtype=torch.float32, device=dev) if curve else empty()
    st = _cabi.lib().mdf_prob_head_fwd_ex(
        v.data_ptr(), w.data_ptr(), hyp.data_ptr(), per_pixel, B, C, D, H, W,
        logits.data_ptr() if want_logits else None, prob.data_ptr() if want_prob else None, depth.data_ptr(),
        conf.data_ptr() if want_confidence else None, n, pad_front, pad_back, upsample,
        _CURVES[curve] if curve else 0, s.data_ptr() if curve else None, int(algo), _stream(v))
    _cabi.check("mdf_prob_head_fwd", st)
    _count(_LAUNCHES_SIMPLE)
    return logits, prob, depth, conf, s


@prob_head.register_fake
def _(x, prob_weight, depth_hypos, curve="", want_logits=False, want_prob=True, want_confidence=False, n=4, pad_front=1,
      pad_back=2, upsample=2, algo=0):
    B, C, D, H, W = x.shape
    e = x.new_empty(0)
    return (x.new_empty((B, D, H, W)) if want_logits else e, x.new_empty((B, D, H, W)) if want_prob else e,
            x.new_empty((B, H, W)), x.new_empty((B, H * upsample, W * upsample)) if want_confidence else e,
            x.new_empty((B, H, W)) if curve else e)


@torch.library.custom_op("mdfnet_b200::depth_regression", mutates_args=(), device_types="cuda")
def depth_regression(prob_volume: Tensor, depth_hypos: Tensor) -> Tensor:
    """regress.py:5-7."""
    p = _prob(prob_volume, "prob_volume")
    B, D, H, W = p.shape
    hyp, per_pixel = _head_hypos(depth_hypos, B, D, H, W)
    out = torch.empty((B, H, W), dtype=torch.float32, device=p.device)
    st = _cabi.lib().mdf_depth_regression_fwd(p.data_ptr(), hyp.data_ptr(), per_pixel, B, D, H, W, out.data_ptr(),
                                              _stream(p))
    _cabi.check("mdf_depth_regression_fwd", st)
    _count(_LAUNCHES_SIMPLE)
    return out


@depth_regression.register_fake
def _(prob_volume, depth_hypos):
    B, _, H, W = prob_volume.shape
    return prob_volume.new_empty((B, H, W))


@torch.library.custom_op("mdfnet_b200::confidence", mutates_args=(), device_types="cuda")
def confidence(prob_volume: Tensor, n: int = 4, pad_front: int = 1, pad_back: int = 2, upsample: int = 1) -> Tensor:
    """regress.py:9-25 (last_confidence=None); upsample=2 folds in core.py:75-77."""
    p = _prob(prob_volume, "prob_volume")
    B, D, H, W = p.shape
    out = torch.empty((B, H * upsample, W * upsample), dtype=torch.float32, device=p.device)
    st = _cabi.lib().mdf_confidence_fwd(p.data_ptr(), B, D, H, W, n, pad_front, pad_back, upsample, out.data_ptr(),
                                        _stream(p))
    _cabi.check("mdf_confidence_fwd", st)
    _count(_LAUNCHES_SIMPLE)
    return out


@confidence.register_fake
def _(prob_volume, n=4, pad_front=1, pad_back=2, upsample=1):
    B, _, H, W = prob_volume.shape
    return prob_volume.new_empty((B, H * upsample, W * upsample))


# ------------------------------------------------------------------- train-mode forward and backward
def _train_common(features, ref_proj, src_projs, depth_hypos, conv_weight, bn_weight, bn_bias, bn_mean, bn_var,
                  fc_weight, fc_bias, groups):
    _check_views(features, src_projs)
    feats = [_f32c(f, "features") for f in features]
    B, C, H, W = feats[0].shape
    hyp, D, per_pixel = _hypos(depth_hypos, B, H, W)
    projs = [_f32c(p, "src_projs") for p in src_projs]
    refp = _f32c(ref_proj, "ref_proj")
    params = [_f32c(t, n) for t, n in ((conv_weight, "conv_weight"), (bn_weight, "bn_weight"), (bn_bias, "bn_bias"),
                                        (bn_mean, "bn_mean"), (bn_var, "bn_var"), (fc_weight, "fc_weight"),
                                        (fc_bias, "fc_bias"))]
    if C != 2 * groups or groups not in (8, 16, 32):
        raise RuntimeError(f"mdfnet_b200: training / autograd supports C == 2*G with G in (8, 16, 32), got C={C}, G={groups}")
    return feats, refp, projs, hyp, per_pixel, params, (B, C, D, H, W)


@torch.library.custom_op("mdfnet_b200::cost_volume_train", mutates_args=(), device_types="cuda")
def cost_volume_train(features: List[Tensor], ref_proj: Tensor, src_projs: List[Tensor], depth_hypos: Tensor,
                      conv_weight: Tensor, bn_weight: Tensor, bn_bias: Tensor, bn_mean: Tensor, bn_var: Tensor,
                      bn_eps: float, fc_weight: Tensor, fc_bias: Tensor, groups: int, training: bool) -> Tuple[Tensor, Tensor]:
    """Cost volume with batch-statistics BatchNorm (training=True) or running statistics (False), computed by the
    kernels whose backward is `cost_volume_bwd`.  Returns (cost volume, (N-1,2) per-view batch mean / unbiased var)."""
    feats, refp, projs, hyp, per_pixel, params, (B, C, D, H, W) = _train_common(
        features, ref_proj, src_projs, depth_hypos, conv_weight, bn_weight, bn_bias, bn_mean, bn_var, fc_weight, fc_bias, groups)
    lib = _cabi.lib()
    N = len(feats)
    dev = feats[0].device
    out = torch.empty((B, groups, D, H, W), dtype=torch.float32, device=dev)
    stats = torch.zeros((N - 1, 2), dtype=torch.float32, device=dev)
    ws = _workspace(lib.mdf_cost_volume_train_workspace_bytes(B, N, C, groups, D, H, W), dev)
    st = lib.mdf_cost_volume_train_fwd(
        _cabi.ptr_array([f.data_ptr() for f in feats]), N, refp.data_ptr(), _cabi.ptr_array([p.data_ptr() for p in projs]),
        hyp.data_ptr(), per_pixel, params[0].data_ptr(), params[1].data_ptr(), params[2].data_ptr(), params[3].data_ptr(),
        params[4].data_ptr(), float(bn_eps), params[5].data_ptr(), params[6].data_ptr(), int(training),
        B, C, groups, D, H, W, out.data_ptr(), stats.data_ptr(), ws.data_ptr(), ws.numel(), _stream(out))
    _cabi.check("mdf_cost_volume_train_fwd", st)
    _count(5 if training else 4)
    return out, stats


@torch.library.custom_op("mdfnet_b200::bn_running_update", mutates_args=("running_mean", "running_var", "num_batches_tracked"),
                         device_types="cuda")
def bn_running_update(batch_stats: Tensor, momentum: float, running_mean: Tensor, running_var: Tensor,
                      num_batches_tracked: Tensor) -> None:
    """The momentum updates of BatchNorm3d's running statistics after a train-mode `cost_volume_train`: one per source view, in
    view order (base.py:50-68 applied per view by homoaggregate.py:40), in place, one launch.  momentum < 0: momentum=None
    (cumulative moving average)."""
    stats = _f32c(batch_stats, "batch_stats")
    if stats.dim() != 2 or stats.shape[1] != 2:
        raise RuntimeError(f"mdfnet_b200: batch_stats must be (N-1,2), got {tuple(stats.shape)}")
    for t, name, dt in ((running_mean, "running_mean", torch.float32), (running_var, "running_var", torch.float32),
                        (num_batches_tracked, "num_batches_tracked", torch.int64)):
        if t.dtype != dt or t.numel() != 1 or not t.is_cuda or not t.is_contiguous():
            raise RuntimeError(f"mdfnet_b200: {name} must be a contiguous CUDA {dt} tensor with one element")
    st = _cabi.lib().mdf_bn_running_update(stats.data_ptr(), stats.shape[0], float(momentum), running_mean.data_ptr(),
                                           running_var.data_ptr(), num_batches_tracked.data_ptr(), _stream(stats))
    _cabi.check("mdf_bn_running_update", st)
    _count(1)


@cost_volume_train.register_fake
def _(features, ref_proj, src_projs, depth_hypos, conv_weight, bn_weight, bn_bias, bn_mean, bn_var, bn_eps,
      fc_weight, fc_bias, groups, training):
    B, _, H, W = features[0].shape
    return features[0].new_empty((B, groups, depth_hypos.shape[1], H, W)), features[0].new_empty((len(features) - 1, 2))


@torch.library.custom_op("mdfnet_b200::cost_volume_bwd", mutates_args=(), device_types="cuda")
def cost_volume_bwd(features: List[Tensor], ref_proj: Tensor, src_projs: List[Tensor], depth_hypos: Tensor,
                    conv_weight: Tensor, bn_weight: Tensor, bn_bias: Tensor, bn_mean: Tensor, bn_var: Tensor,
                    bn_eps: float, fc_weight: Tensor, fc_bias: Tensor, groups: int, training: bool,
                    cost_volume: Tensor, grad_out: Tensor, batch_stats: Tensor) -> Tuple[List[Tensor], Tensor]:
    """Gradients of `cost_volume_train`: ([d features[i]], (4+G,) = d bn.weight, d bn.bias, d fc.weight, d fc.bias, d conv.weight)."""
    feats, refp, projs, hyp, per_pixel, params, (B, C, D, H, W) = _train_common(
        features, ref_proj, src_projs, depth_hypos, conv_weight, bn_weight, bn_bias, bn_mean, bn_var, fc_weight, fc_bias, groups)
    lib = _cabi.lib()
    N = len(feats)
    dev = feats[0].device
    cv, go = _f32c(cost_volume, "cost_volume"), _f32c(grad_out, "grad_out")
    if cv.shape != (B, groups, D, H, W) or go.shape != cv.shape:
        raise RuntimeError("mdfnet_b200: cost_volume / grad_out must be (B,G,D,H,W)")
    gfeats = [torch.empty_like(f) for f in feats]
    gparams = torch.empty(4 + groups, dtype=torch.float32, device=dev)
    ws = _workspace(lib.mdf_cost_volume_train_workspace_bytes(B, N, C, groups, D, H, W), dev)
    st = lib.mdf_cost_volume_bwd(
        _cabi.ptr_array([f.data_ptr() for f in feats]), N, refp.data_ptr(), _cabi.ptr_array([p.data_ptr() for p in projs]),
        hyp.data_ptr(), per_pixel, params[0].data_ptr(), params[1].data_ptr(), params[2].data_ptr(), params[3].data_ptr(),
        params[4].data_ptr(), float(bn_eps), params[5].data_ptr(), params[6].data_ptr(), int(training),
        B, C, groups, D, H, W, cv.data_ptr(), go.data_ptr(),
        _f32c(batch_stats, "batch_stats").data_ptr() if training and batch_stats.numel() else None,
        _cabi.ptr_array([g.data_ptr() for g in gfeats]), gparams.data_ptr(), ws.data_ptr(), ws.numel(), _stream(cv))
    _cabi.check("mdf_cost_volume_bwd", st)
    _count(8)           # setup, layout pass, fold, staged gather (phase 1a), per-element pass (1b), plane sweep, finish, parameter gradients
    return gfeats, gparams


@cost_volume_bwd.register_fake
def _(features, ref_proj, src_projs, depth_hypos, conv_weight, bn_weight, bn_bias, bn_mean, bn_var, bn_eps,
      fc_weight, fc_bias, groups, training, cost_volume, grad_out, batch_stats):
    return [torch.empty_like(f) for f in features], features[0].new_empty(4 + groups)


# ------------------------------------------------------------------ next-stage hypotheses (HyposByFit)
@torch.library.custom_op("mdfnet_b200::hypos_fit", mutates_args=(), device_types="cuda")
def hypos_fit(prob_volume: Tensor, depth_hypos: Tensor, depth: Tensor, curve: str) -> Tensor:
    """Fitted scale s (B,H,W) of every pixel's probability column: depthhypos.py:78-125 ('laplace'), :169-215 ('gauss1')."""
    if curve not in _FIT_CURVES:
        raise RuntimeError(f"mdfnet_b200: curve must be one of {sorted(_FIT_CURVES)}, got {curve!r}")
    p = _prob(prob_volume, "prob_volume")
    B, D, H, W = p.shape
    hyp, per_pixel = _head_hypos(depth_hypos, B, D, H, W)
    d = _f32c(depth, "depth")
    if d.shape != (B, H, W):
        raise RuntimeError(f"mdfnet_b200: depth must be (B,H,W) = {(B, H, W)}, got {tuple(d.shape)}")
    s = torch.empty((B, H, W), dtype=torch.float32, device=p.device)
    st = _cabi.lib().mdf_hypos_fit_fwd(p.data_ptr(), hyp.data_ptr(), per_pixel, d.data_ptr(), _FIT_CURVES[curve], B, D, H, W,
                                       s.data_ptr(), _stream(p))
    _cabi.check("mdf_hypos_fit_fwd", st)
    _count(_LAUNCHES_SIMPLE)
    return s


@hypos_fit.register_fake
def _(prob_volume, depth_hypos, depth, curve):
    B, _, H, W = prob_volume.shape
    return prob_volume.new_empty((B, H, W))


@torch.library.custom_op("mdfnet_b200::hypos_generate", mutates_args=(), device_types="cuda")
def hypos_generate(depth: Tensor, s: Tensor, depth_range: Tensor, curve: str, prob_thresh: float, ndepths: int,
                   upsample: bool) -> Tensor:
    """Hypotheses (B,ndepths,2H|H,2W|W) from the fitted scale: depthhypos.py:48-76."""
    if curve not in _FIT_CURVES:
        raise RuntimeError(f"mdfnet_b200: curve must be one of {sorted(_FIT_CURVES)}, got {curve!r}")
    d, sv = _f32c(depth, "depth"), _f32c(s, "s")
    B, H, W = d.shape
    rng = _f32c(depth_range.float(), "depth_range")
    if sv.shape != d.shape or rng.shape != (B, 2):
        raise RuntimeError("mdfnet_b200: s must match depth (B,H,W) and depth_range must be (B,2)")
    f = 2 if upsample else 1
    out = torch.empty((B, ndepths, H * f, W * f), dtype=torch.float32, device=d.device)
    st = _cabi.lib().mdf_hypos_generate_fwd(d.data_ptr(), sv.data_ptr(), rng.data_ptr(), _FIT_CURVES[curve], float(prob_thresh),
                                            int(upsample), B, H, W, int(ndepths), out.data_ptr(), _stream(d))
    _cabi.check("mdf_hypos_generate_fwd", st)
    _count(_LAUNCHES_SIMPLE)
    return out


@hypos_generate.register_fake
def _(depth, s, depth_range, curve, prob_thresh, ndepths, upsample):
    B, H, W = depth.shape
    f = 2 if upsample else 1
    return depth.new_empty((B, ndepths, H * f, W * f))


# ------------------------------------------------------------- post-processing: geometric-consistency filter
def geo_filter(ref_depth: Tensor, ref_intrinsics: Tensor, ref_extrinsics: Tensor, src_depths: Sequence[Tensor],
               src_intrinsics: Tensor, src_extrinsics: Tensor, confidence=None, photo_threshold: float = 0.8,
               nconditions: int = 5, thre1: float = 4.0, thre2: float = 1300.0, per_source: bool = False) -> dict:
    """One reference view of tools/filter/dynamic_filter_gpu.py:57-100,161-237 in one launch.  Returns a dict with
    depth_averaged (H,W) float32 and geo / photo / final (H,W) bool; with per_source=True also bits (S,H,W) int16 (bit
    i-2 = dynamic mask of threshold i) and depth_reprojected (S,H,W).  (Not a torch.library op: it returns masks of
    several dtypes and is only ever called eagerly, from the post-processing script.)"""
    d = _f32c(ref_depth, "ref_depth")
    if d.dim() != 2:
        raise RuntimeError(f"mdfnet_b200: ref_depth must be (H,W), got {tuple(d.shape)}")
    H, W = d.shape
    srcs = [_f32c(s, "src_depths") for s in src_depths]
    S = len(srcs)
    if any(s.shape != d.shape for s in srcs):
        raise RuntimeError("mdfnet_b200: every source depth map must have the reference map's (H,W)")
    dev = d.device
    K, E = _f32c(ref_intrinsics, "ref_intrinsics"), _f32c(ref_extrinsics, "ref_extrinsics")
    sK = _f32c(src_intrinsics, "src_intrinsics") if S else torch.empty(0, device=dev)
    sE = _f32c(src_extrinsics, "src_extrinsics") if S else torch.empty(0, device=dev)
    if K.numel() != 9 or E.numel() != 16 or sK.numel() != 9 * S or sE.numel() != 16 * S:
        raise RuntimeError("mdfnet_b200: intrinsics must be 3x3 / (S,3,3) and extrinsics 4x4 / (S,4,4)")
    conf = _f32c(confidence, "confidence") if confidence is not None else None
    lib = _cabi.lib()
    out = dict(depth_averaged=torch.empty((H, W), dtype=torch.float32, device=dev),
               geo=torch.empty((H, W), dtype=torch.bool, device=dev), photo=torch.empty((H, W), dtype=torch.bool, device=dev),
               final=torch.empty((H, W), dtype=torch.bool, device=dev))
    if per_source:
        out["bits"] = torch.empty((S, H, W), dtype=torch.int16, device=dev)
        out["depth_reprojected"] = torch.empty((S, H, W), dtype=torch.float32, device=dev)
    ws = _workspace(lib.mdf_geo_filter_workspace_bytes(S), dev)
    st = lib.mdf_geo_filter_fwd(
        d.data_ptr(), K.data_ptr(), E.data_ptr(), _cabi.ptr_array([s.data_ptr() for s in srcs]) if S else None,
        sK.data_ptr() if S else None, sE.data_ptr() if S else None, S, H, W, conf.data_ptr() if conf is not None else None,
        float(photo_threshold), int(nconditions), float(thre1), float(thre2),
        out["bits"].data_ptr() if per_source else None, out["depth_reprojected"].data_ptr() if per_source else None,
        out["depth_averaged"].data_ptr(), out["geo"].data_ptr(), out["photo"].data_ptr(), out["final"].data_ptr(),
        ws.data_ptr(), ws.numel(), _stream(d))
    _cabi.check("mdf_geo_filter_fwd", st)
    _count(2)
    return out
