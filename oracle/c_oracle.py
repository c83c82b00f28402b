"""ctypes/numpy front-end of the C oracle (oracle/mdf_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(mdf_net_b200/) never imports it.

Two precisions are built from the same source: "f32" is the oracle proper (the
reference computes in float32), "f64" evaluates the same formulae in double and
is used only to measure the float32 noise floor of the reference itself.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_LIBS: dict = {}


def build(force: bool = False) -> None:
    """Compile the oracle with the committed Makefile (gcc, seconds)."""
    targets = [os.path.join(_BUILD, f"libmdf_oracle_{p}.so") for p in ("f32", "f64")]
    src = os.path.join(_HERE, "mdf_oracle.c")
    fresh = all(os.path.exists(t) and os.path.getmtime(t) >= os.path.getmtime(src) for t in targets)
    if fresh and not force:
        return
    for extra in ([], ["OMP="]):  # second attempt: no OpenMP runtime in the image
        r = subprocess.run(["make", "-C", _HERE, "-B"] + extra, capture_output=True, text=True)
        if r.returncode == 0:
            return
    raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)


def _lib(prec: str):
    if prec not in _LIBS:
        path = os.path.join(_BUILD, f"libmdf_oracle_{prec}.so")
        if not os.path.exists(path):
            build()
        _LIBS[prec] = ctypes.CDLL(path)
    return _LIBS[prec]


def _dt(prec: str):
    return np.float32 if prec == "f32" else np.float64


def _arr(a, prec):
    return np.ascontiguousarray(np.asarray(a), dtype=_dt(prec))


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _ptr_array(arrs):
    return (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def _real(prec):
    return ctypes.c_float if prec == "f32" else ctypes.c_double


def _check(rc, what):
    if rc != 0:
        raise ValueError(f"{what}: oracle returned {rc}")


def _hypos(depth_hypos, prec, B, D, H, W):
    h = _arr(depth_hypos, prec)
    if h.shape == (B, D, H, W) and not (H == 1 and W == 1):
        return h, 1
    if h.size == B * D:
        return h.reshape(B, D), 0
    raise ValueError(f"depth_hypos shape {h.shape} is neither (B,D,1,1) nor (B,D,H,W)")


def compose_proj(src_proj, ref_proj, prec="f32"):
    """(B,4,4),(B,4,4) -> (B,12): rot row-major (9) then trans (3).  base.py:98-100."""
    s, r = _arr(src_proj, prec), _arr(ref_proj, prec)
    B = s.shape[0]
    out = np.empty((B, 12), _dt(prec))
    fn = getattr(_lib(prec), f"mdf_oracle_compose_proj_{prec}")
    _check(fn(_ptr(s), _ptr(r), B, _ptr(out)), "compose_proj")
    return out


def homo_warp(src_fea, depth_hypos, rot_trans=None, src_proj=None, ref_proj=None, prec="f32"):
    """homo_warping (base.py:85-126).  Pass either rot_trans (B,12) or the two projections."""
    f = _arr(src_fea, prec)
    B, C, H, W = f.shape
    D = np.asarray(depth_hypos).shape[1]
    h, pp = _hypos(depth_hypos, prec, B, D, H, W)
    rt = _arr(rot_trans, prec) if rot_trans is not None else compose_proj(src_proj, ref_proj, prec)
    out = np.empty((B, C, D, H, W), _dt(prec))
    fn = getattr(_lib(prec), f"mdf_oracle_homo_warp_{prec}")
    _check(fn(_ptr(f), _ptr(rt), _ptr(h), pp, B, C, D, H, W, _ptr(out)), "homo_warp")
    return out


def sample_positions(rot_trans, depth_hypos, H, W, prec="f32"):
    """(ix, iy), each (D,H,W): where grid_sample reads for every reference pixel and hypothesis of ONE
    batch item (base.py:102-119 + ATen unnormalize).  rot_trans: 12 numbers; depth_hypos (D,) or (D,H,W)."""
    rt = _arr(rot_trans, prec).reshape(12)
    h = _arr(depth_hypos, prec)
    D = h.shape[0]
    per_pixel = 1 if h.ndim == 3 else 0
    ix = np.empty((D, H, W), _dt(prec)); iy = np.empty((D, H, W), _dt(prec))
    fn = getattr(_lib(prec), f"mdf_oracle_sample_positions_{prec}")
    _check(fn(_ptr(rt), _ptr(h), per_pixel, D, H, W, _ptr(ix), _ptr(iy)), "sample_positions")
    return ix, iy


def _rot_trans_list(ref_proj, src_projs, rot_trans, prec):
    if rot_trans is not None:
        return [_arr(r, prec) for r in rot_trans]
    return [compose_proj(s, ref_proj, prec) for s in src_projs]


def vector_aggregate(features: Sequence, depth_hypos, params: dict, G: int,
                     ref_proj=None, src_projs=None, rot_trans=None, prec="f32", rows=None):
    """VectorAggregate.forward in eval mode (homoaggregate.py:25-46).

    params: {"cw": (G,), "bn_weight", "bn_bias", "bn_mean", "bn_var", "bn_eps", "fc_weight", "fc_bias"}
    rows=(y0, y1): evaluate only those rows (bench.py's bounded sample); returns (B,G,D,y1-y0,W).
    """
    feats = [_arr(f, prec) for f in features]
    B, C, H, W = feats[0].shape
    N = len(feats)
    D = np.asarray(depth_hypos).shape[1]
    h, pp = _hypos(depth_hypos, prec, B, D, H, W)
    rts = _rot_trans_list(ref_proj, src_projs, rot_trans, prec)
    cw = _arr(np.asarray(params["cw"]).reshape(-1), prec)
    out = np.empty((B, G, D, H, W), _dt(prec))
    R = _real(prec)
    y0, y1 = (0, H) if rows is None else (int(rows[0]), int(rows[1]))
    fn = getattr(_lib(prec), f"mdf_oracle_vector_aggregate_rows_{prec}")
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                   ctypes.c_void_p, R, R, R, R, R, R, R] + [ctypes.c_int] * 8 + [ctypes.c_void_p]
    rc = fn(_ptr_array(feats), _ptr_array(rts), N, _ptr(h), pp, _ptr(cw),
            float(params["bn_weight"]), float(params["bn_bias"]), float(params["bn_mean"]),
            float(params["bn_var"]), float(params.get("bn_eps", 1e-5)),
            float(params["fc_weight"]), float(params["fc_bias"]),
            B, C, G, D, H, W, y0, y1, _ptr(out))
    _check(rc, "vector_aggregate")
    return out if rows is None else out[:, :, :, y0:y1]


def variance_aggregate(features: Sequence, depth_hypos, ref_proj=None, src_projs=None, rot_trans=None, prec="f32"):
    """homo_aggregate_by_variance (homoaggregate.py:49-69)."""
    feats = [_arr(f, prec) for f in features]
    B, C, H, W = feats[0].shape
    D = np.asarray(depth_hypos).shape[1]
    h, pp = _hypos(depth_hypos, prec, B, D, H, W)
    rts = _rot_trans_list(ref_proj, src_projs, rot_trans, prec)
    out = np.empty((B, C, D, H, W), _dt(prec))
    fn = getattr(_lib(prec), f"mdf_oracle_variance_aggregate_{prec}")
    _check(fn(_ptr_array(feats), _ptr_array(rts), len(feats), _ptr(h), pp, B, C, D, H, W, _ptr(out)),
           "variance_aggregate")
    return out


def softmax_depth(logits, prec="f32"):
    """F.softmax(x, dim=1) tail of the regulariser (regular.py:69,133)."""
    x = _arr(logits, prec)
    B, D, H, W = x.shape
    out = np.empty_like(x)
    fn = getattr(_lib(prec), f"mdf_oracle_softmax_depth_{prec}")
    _check(fn(_ptr(x), B, D, H, W, _ptr(out)), "softmax_depth")
    return out


def prob_conv(x, weight, prec="f32"):
    """x = self.prob(x).squeeze(1): Conv3d(c0, 1, 3, padding=1, bias=False) (regular.py:43,67 / :110,130) -> logits (B,D,H,W)."""
    v = _arr(x, prec)
    B, C, D, H, W = v.shape
    w = _arr(weight, prec).reshape(C, 3, 3, 3)
    out = np.empty((B, D, H, W), _dt(prec))
    fn = getattr(_lib(prec), f"mdf_oracle_prob_conv_{prec}")
    _check(fn(_ptr(v), _ptr(w), B, C, D, H, W, _ptr(out)), "prob_conv")
    return out


def depth_regression(prob_volume, depth_hypos, prec="f32"):
    """regress.py:5-7."""
    p = _arr(prob_volume, prec)
    B, D, H, W = p.shape
    h, pp = _hypos(depth_hypos, prec, B, D, H, W)
    out = np.empty((B, H, W), _dt(prec))
    fn = getattr(_lib(prec), f"mdf_oracle_depth_regression_{prec}")
    _check(fn(_ptr(p), _ptr(h), pp, B, D, H, W, _ptr(out)), "depth_regression")
    return out


def confidence_regress(prob_volume, n=4, pad=(0, 0, 0, 0, 1, 2), upsample=1, prec="f32"):
    """regress.py:9-25 with last_confidence=None; upsample=2 adds core.py:75-77."""
    if tuple(pad[:4]) != (0, 0, 0, 0):
        raise ValueError("only depth padding is meaningful here")
    p = _arr(prob_volume, prec)
    B, D, H, W = p.shape
    out = np.empty((B, H * upsample, W * upsample), _dt(prec))
    fn = getattr(_lib(prec), f"mdf_oracle_confidence_{prec}")
    _check(fn(_ptr(p), B, D, H, W, int(n), int(pad[4]), int(pad[5]), int(upsample), _ptr(out)), "confidence")
    return out


def num_threads() -> int:
    """Threads the OpenMP build uses (1 when built with OMP=)."""
    lib = _lib("f32")
    try:
        fn = lib.omp_get_max_threads
    except AttributeError:
        return 1
    fn.restype = ctypes.c_int
    return int(fn())


def set_num_threads(n: int) -> int:
    """Use `n` OpenMP threads from now on (torchrun exports OMP_NUM_THREADS=1; the CPU baseline wants them all).
    Returns the number in effect (1 when built without OpenMP)."""
    lib = _lib("f32")
    try:
        fn = lib.omp_set_num_threads
    except AttributeError:
        return 1
    fn.argtypes = [ctypes.c_int]
    fn(int(max(1, n)))
    return num_threads()


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


_FIT_MODES = {"gauss1": 1, "laplace": 2, "gauss0": 3}


def hypos_fit(prob_volume, depth_hypos, depth, curve, prec="f32"):
    """Per-pixel curve fit of HyposByFit (depthhypos.py:78-125 'laplace', :127-167 'gauss0', :169-215 'gauss1') -> s (B,H,W)."""
    p = _arr(prob_volume, prec)
    B, D, H, W = p.shape
    h, pp = _hypos(depth_hypos, prec, B, D, H, W)
    d = _arr(depth, prec)
    out = np.empty((B, H, W), _dt(prec))
    fn = getattr(_lib(prec), f"mdf_oracle_hypos_fit_{prec}")
    _check(fn(_ptr(p), _ptr(h), pp, _ptr(d), _FIT_MODES[curve], B, D, H, W, _ptr(out)), "hypos_fit")
    return out


def hypos_generate(depth, s, depth_range, curve, prob_thresh, ndepths, upsample=True, prec="f32"):
    """Next-stage hypotheses from the fitted scale (depthhypos.py:48-76) -> (B, ndepths, 2H|H, 2W|W)."""
    d, sv = _arr(depth, prec), _arr(s, prec)
    B, H, W = d.shape
    dr = _arr(depth_range, prec).reshape(B, 2)
    out = np.empty((B, ndepths, H * (2 if upsample else 1), W * (2 if upsample else 1)), _dt(prec))
    fn = getattr(_lib(prec), f"mdf_oracle_hypos_generate_{prec}")
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, _real(prec)] + [ctypes.c_int] * 5 + [ctypes.c_void_p]
    _check(fn(_ptr(d), _ptr(sv), _ptr(dr), _FIT_MODES[curve], float(prob_thresh), int(bool(upsample)), B, H, W, int(ndepths),
              _ptr(out)), "hypos_generate")
    return out


def geo_filter(ref_depth, ref_K, ref_E, src_depths, src_K, src_E, confidence=None, photo_threshold=0.8, nconditions=5,
               thre1=4.0, thre2=1300.0, prec="f32"):
    """Geometric-consistency filter of one reference view (tools/filter/dynamic_filter_gpu.py:57-100,161-237).
    Returns dict(bits (S,H,W) uint16, depth_reprojected (S,H,W), depth_averaged (H,W), geo, photo, final (H,W) uint8)."""
    d = _arr(ref_depth, prec)
    H, W = d.shape
    srcs = [_arr(x, prec) for x in src_depths]
    S = len(srcs)
    K, E = _arr(ref_K, prec).reshape(3, 3), _arr(ref_E, prec).reshape(4, 4)
    sK, sE = _arr(src_K, prec).reshape(S, 3, 3), _arr(src_E, prec).reshape(S, 4, 4)
    conf = _arr(confidence, prec) if confidence is not None else None
    out = dict(bits=np.zeros((S, H, W), np.uint16), depth_reprojected=np.zeros((S, H, W), _dt(prec)),
               depth_averaged=np.zeros((H, W), _dt(prec)), geo=np.zeros((H, W), np.uint8), photo=np.zeros((H, W), np.uint8),
               final=np.zeros((H, W), np.uint8))
    fn = getattr(_lib(prec), f"mdf_oracle_geo_filter_{prec}")
    R = _real(prec)
    fn.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int] * 3 + [ctypes.c_void_p, R, ctypes.c_int, R, R] + [ctypes.c_void_p] * 6
    _check(fn(_ptr(d), _ptr(K), _ptr(E), _ptr_array(srcs), _ptr(sK), _ptr(sE), S, H, W, _ptr(conf) if conf is not None else None,
              float(photo_threshold), int(nconditions), float(thre1), float(thre2), _ptr(out["bits"]),
              _ptr(out["depth_reprojected"]), _ptr(out["depth_averaged"]), _ptr(out["geo"]), _ptr(out["photo"]),
              _ptr(out["final"])), "geo_filter")
    return out
