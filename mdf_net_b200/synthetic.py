"""Seeded synthetic inputs shaped like the reference's DTU / Tanks / BlendedMVS samples.

The reference ships no data and its checkpoints are absent from the checkout, so every
parity test, the golden fixtures and bench.py draw their inputs from here.  Shapes and value
ranges follow the reference loaders: images in [0,1) (tools/data_io.py:103-107), DTU depth
range [425, 935] mm (load/dtueval.py:47), per-stage feature widths C=(64,32,16), hypotheses
D=(48,24,8), groups G=(32,16,8) (config.py:196,199,205).

numpy only (PCG64 streams are stable across machines), float32 throughout.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

STAGE_CHANNELS = (64, 32, 16)   # config.py:196 + backbone.py:59-66
STAGE_DEPTHS = (48, 24, 8)      # config.py:199
STAGE_GROUPS = (32, 16, 8)      # config.py:205
DTU_DEPTH_RANGE = (425.0, 935.0)  # load/dtueval.py:47


def stage_shapes(h0: int, w0: int) -> List[Tuple[int, int]]:
    """Feature-map sizes of the three cost-volume stages (1/8, 1/4, 1/2 of the image)."""
    return [(h0 // 8, w0 // 8), (h0 // 4, w0 // 4), (h0 // 2, w0 // 2)]


def _rot(axis: str, a: float) -> np.ndarray:
    c, s = math.cos(a), math.sin(a)
    if axis == "x":
        return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], np.float64)
    if axis == "y":
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float64)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], np.float64)


def camera_rig(batch: int, nviews: int, h0: int, w0: int, seed: int = 1, general: bool = True,
               focus: float = 700.0) -> Tuple[np.ndarray, np.ndarray]:
    """DTU-like intrinsics (B,N,3,3) and extrinsics (B,N,4,4) at image size h0 x w0.

    View 0 is the reference.  Source view v orbits the point `focus` mm in front of the
    reference camera by +-0.1*ceil(v/2) rad about y; with `general` a small roll,
    pitch and vertical baseline are added so epipolar lines are not axis aligned.
    """
    rng = np.random.default_rng(seed)
    K = np.zeros((batch, nviews, 3, 3), np.float64)
    E = np.zeros((batch, nviews, 4, 4), np.float64)
    for b in range(batch):
        # world -> reference camera: a mild general pose so inverse(ref_proj) is not trivial
        Rw = _rot("y", 0.2 + 0.01 * b) @ _rot("x", -0.1)
        tw = np.array([100.0, -50.0, 30.0])
        for v in range(nviews):
            fx = 2892.33 * w0 / 1600.0 * (1.0 + 0.002 * v)
            fy = 2883.18 * w0 / 1600.0 * (1.0 + 0.002 * v)
            K[b, v] = [[fx, 0, w0 / 2.0 + 0.3 * v], [0, fy, h0 / 2.0 - 0.2 * v], [0, 0, 1]]
            if v == 0:
                Rrel, trel = np.eye(3), np.zeros(3)
            else:
                a = 0.1 * math.ceil(v / 2) * (1 if v % 2 else -1)
                Rrel = _rot("y", a)
                trel = np.array([-focus * math.sin(a), 0.0, focus * (1 - math.cos(a))])
                if general:
                    roll, pitch = rng.uniform(-0.03, 0.03), rng.uniform(-0.03, 0.03)
                    Rg = _rot("z", roll) @ _rot("x", pitch)
                    # keep the focus point fixed under the extra rotation
                    p = np.array([0.0, 0.0, focus])
                    Rrel = Rg @ Rrel
                    trel = Rg @ trel + (p - Rg @ p) + np.array([0.0, rng.uniform(-25, 25), 0.0])
            # X_src = Rrel (Rw Xw + tw) + trel
            E[b, v, :3, :3] = Rrel @ Rw
            E[b, v, :3, 3] = Rrel @ tw + trel
            E[b, v, 3, 3] = 1.0
    return K.astype(np.float32), E.astype(np.float32)


def smooth_features(batch: int, nviews: int, channels: int, h: int, w: int, seed: int = 2,
                    scale: float = 2.0) -> List[np.ndarray]:
    """N feature maps (B,C,H,W): 3x3 box-filtered N(0,1) times `scale`."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(nviews):
        x = rng.standard_normal((batch, channels, h, w), dtype=np.float32)
        p = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
        acc = np.zeros_like(x)
        for dy in range(3):
            for dx in range(3):
                acc += p[:, :, dy:dy + h, dx:dx + w]
        out.append((acc * np.float32(scale / 9.0)).astype(np.float32))
    return out


def uniform_hypos(batch: int, ndepths: int, dmin: float = DTU_DEPTH_RANGE[0],
                  dmax: float = DTU_DEPTH_RANGE[1]) -> np.ndarray:
    """Stage-0 hypotheses (B,D,1,1), as depthhypos.py:31-38 builds them."""
    interval = np.float32((np.float32(dmax) - np.float32(dmin)) / np.float32(ndepths - 1))
    h = np.float32(dmin) + np.arange(ndepths, dtype=np.float32) * interval
    return np.broadcast_to(h.reshape(1, ndepths, 1, 1), (batch, ndepths, 1, 1)).astype(np.float32).copy()


def pixel_hypos(batch: int, ndepths: int, h: int, w: int, seed: int = 3,
                dmin: float = DTU_DEPTH_RANGE[0], dmax: float = DTU_DEPTH_RANGE[1],
                max_rel_range: float = 0.06, smooth: bool = True) -> np.ndarray:
    """Stage-1/2 hypotheses (B,D,H,W): d0 - r/2 + k*r/(D-1), clamped (depthhypos.py:58-74)."""
    rng = np.random.default_rng(seed)
    d0 = rng.uniform(500.0, 800.0, (batch, 1, h, w)).astype(np.float32)
    r = rng.uniform(0.2, 1.0, (batch, 1, h, w)).astype(np.float32) * np.float32(max_rel_range * (dmax - dmin))
    if smooth:  # neighbouring pixels see similar surfaces
        for arr in (d0, r):
            p = np.pad(arr, ((0, 0), (0, 0), (2, 2), (2, 2)), mode="edge")
            acc = np.zeros_like(arr)
            for dy in range(5):
                for dx in range(5):
                    acc += p[:, :, dy:dy + h, dx:dx + w]
            arr[...] = acc / np.float32(25.0)
    k = np.arange(ndepths, dtype=np.float32).reshape(1, ndepths, 1, 1)
    hyp = d0 - np.float32(0.5) * r + k * (r / np.float32(ndepths - 1))
    return np.clip(hyp, np.float32(dmin), np.float32(dmax)).astype(np.float32)


def _upsample2x_bilinear(a: np.ndarray) -> np.ndarray:
    """(B,1,h,w) -> (B,1,2h,2w), F.interpolate(scale_factor=2, mode='bilinear', align_corners=False) semantics:
    the way the reference carries depth / interval maps to the next stage (depthhypos.py:42-52)."""
    def up(x, axis):
        n = x.shape[axis]
        pos = (np.arange(2 * n, dtype=np.float32) + 0.5) / 2.0 - 0.5
        pos = np.clip(pos, 0.0, n - 1)
        i0 = np.floor(pos).astype(np.int64)
        i1 = np.minimum(i0 + 1, n - 1)
        w1 = (pos - i0).astype(np.float32)
        shape = [1] * x.ndim
        shape[axis] = 2 * n
        w1 = w1.reshape(shape)
        return np.take(x, i0, axis=axis) * (1.0 - w1) + np.take(x, i1, axis=axis) * w1
    return up(up(a, 2), 3).astype(np.float32)


def scene_depth(batch: int, h: int, w: int, seed: int = 6) -> np.ndarray:
    """A DTU-like synthetic scene as a depth map (B,1,h,w) in mm: a slanted background plane around 780 mm
    and a few ellipsoidal objects (520-640 mm at the centre, bowl shaped, 50-200 mm silhouette jumps),
    with fine surface relief.  Evaluated analytically in normalised image coordinates, so every
    resolution sees the same scene."""
    rng = np.random.default_rng(seed)
    v, u = np.meshgrid((np.arange(h, dtype=np.float32) + 0.5) / h, (np.arange(w, dtype=np.float32) + 0.5) / w, indexing="ij")
    out = np.empty((batch, 1, h, w), np.float32)
    for b in range(batch):
        z = 780.0 + 60.0 * (u - 0.5) - 40.0 * (v - 0.5) + 3.0 * np.sin(40.0 * u + b) * np.cos(35.0 * v)
        for _ in range(4):
            cu, cv = rng.uniform(0.2, 0.8), rng.uniform(0.25, 0.75)
            ru, rv = rng.uniform(0.10, 0.28), rng.uniform(0.12, 0.30)
            peak = rng.uniform(520.0, 640.0)
            rr = ((u - cu) / ru) ** 2 + ((v - cv) / rv) ** 2
            obj = peak + 90.0 * rr + 2.0 * np.sin(90.0 * u) * np.sin(80.0 * v)
            z = np.where(rr < 1.0, np.minimum(z, obj), z)
        out[b, 0] = z
    return out


def scene_hypos(batch: int, ndepths: int, h: int, w: int, seed: int = 6, dmin: float = DTU_DEPTH_RANGE[0],
                dmax: float = DTU_DEPTH_RANGE[1], range_mm: Tuple[float, float] = (8.0, 30.0)) -> np.ndarray:
    """Stage-1/2 hypotheses (B,D,H,W) the way the reference forms them (depthhypos.py:40-74): the depth and
    the per-pixel search range of the previous (half resolution) stage are bilinearly upsampled x2, then
    d - r/2 + k*r/(D-1), clamped to the depth range.  The previous stage's depth is `scene_depth` at half
    resolution; the range varies smoothly between range_mm[0] and range_mm[1]."""
    rng = np.random.default_rng(seed + 1)
    hh, hw = (h + 1) // 2, (w + 1) // 2
    d0 = _upsample2x_bilinear(scene_depth(batch, hh, hw, seed))[:, :, :h, :w]
    v, u = np.meshgrid(np.linspace(0, 1, hh, dtype=np.float32), np.linspace(0, 1, hw, dtype=np.float32), indexing="ij")
    ph = rng.uniform(0, 6.28, 4)
    t = 0.5 + 0.25 * np.sin(5.0 * u + ph[0]) * np.cos(4.0 * v + ph[1]) + 0.25 * np.sin(9.0 * v + ph[2]) * np.cos(7.0 * u + ph[3])
    r = (range_mm[0] + (range_mm[1] - range_mm[0]) * t).astype(np.float32)
    r = _upsample2x_bilinear(np.broadcast_to(r, (batch, 1, hh, hw)).copy())[:, :, :h, :w]
    k = np.arange(ndepths, dtype=np.float32).reshape(1, ndepths, 1, 1)
    hyp = d0 - np.float32(0.5) * r + k * (r / np.float32(ndepths - 1))
    return np.clip(hyp, np.float32(dmin), np.float32(dmax)).astype(np.float32)


def depth_weight_params(groups: int, seed: int = 4) -> dict:
    """Non-degenerate depth_weight parameters (homoaggregate.py:16-20)."""
    rng = np.random.default_rng(seed)
    return {
        "cw": (rng.standard_normal(groups) * 0.5).astype(np.float32),
        "bn_weight": np.float32(1.0 + 0.2 * rng.standard_normal()),
        "bn_bias": np.float32(0.1 * rng.standard_normal()),
        "bn_mean": np.float32(0.3),
        "bn_var": np.float32(0.7),
        "bn_eps": 1e-5,
        "fc_weight": np.float32(0.8 + 0.3 * rng.standard_normal()),
        "fc_bias": np.float32(0.2 * rng.standard_normal()),
    }


def regulariser_logits(batch: int, ndepths: int, h: int, w: int, seed: int = 5, peak: float = 4.0) -> np.ndarray:
    """Stand-in for the 3-D CNN output (B,D,H,W): noise plus one bump per pixel."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((batch, ndepths, h, w), dtype=np.float32)
    centre = rng.uniform(0, ndepths - 1, (batch, 1, h, w)).astype(np.float32)
    k = np.arange(ndepths, dtype=np.float32).reshape(1, ndepths, 1, 1)
    width = np.float32(max(1.0, ndepths / 8.0))
    return (x + np.float32(peak) * np.exp(-0.5 * ((k - centre) / width) ** 2)).astype(np.float32)


def scene_logits(batch: int, stage: int, h: int, w: int, seed: int = 6, dmin: float = DTU_DEPTH_RANGE[0],
                 dmax: float = DTU_DEPTH_RANGE[1], peak: float = 9.0) -> np.ndarray:
    """Stand-in for the 3-D CNN output (B,D,H,W) of `stage` that makes the coarse-to-fine CHAIN behave like a real scene:
    one bump per pixel plus noise, placed so that stage 0 regresses to `scene_depth` over the uniform hypotheses and the
    later stages regress to a smooth offset around the middle of whatever hypotheses HyposByFit hands them.  The bump
    width varies smoothly over the image, so the fitted search ranges of the next stage do too (8-30 mm, like
    `scene_hypos`).  Used where the hypotheses of stages 1-2 are produced on the device (bench.py's end-to-end leg)."""
    rng = np.random.default_rng(seed + 100 + stage)
    D = STAGE_DEPTHS[stage]
    v, u = np.meshgrid(np.linspace(0, 1, h, dtype=np.float32), np.linspace(0, 1, w, dtype=np.float32), indexing="ij")
    ph = rng.uniform(0, 6.28, 4)
    t = 0.5 + 0.25 * np.sin(5.0 * u + ph[0]) * np.cos(4.0 * v + ph[1]) + 0.25 * np.sin(9.0 * v + ph[2]) * np.cos(7.0 * u + ph[3])
    if stage == 0:
        interval = (dmax - dmin) / (D - 1)
        centre = (scene_depth(batch, h, w, seed) - np.float32(dmin)) / np.float32(interval)        # (B,1,h,w) in planes
        width = (0.6 + 1.0 * t)[None, None]
    else:
        centre = np.broadcast_to(((D - 1) / 2.0 + (D / 8.0) * (2.0 * t - 1.0))[None, None], (batch, 1, h, w))
        width = (np.float32(D) / 16.0 + 0.6 * t)[None, None]
    k = np.arange(D, dtype=np.float32).reshape(1, D, 1, 1)
    x = np.float32(0.3) * rng.standard_normal((batch, D, h, w), dtype=np.float32)
    return (x + np.float32(peak) * np.exp(-0.5 * ((k - centre) / width) ** 2)).astype(np.float32)


def projection_matrices(K: np.ndarray, E: np.ndarray, level_div: float = 1.0) -> np.ndarray:
    """(B,N,4,4) float32 P = [K/level_div (rows 0-1) @ E[:3,:4]; E[3]]  -- what scale_cam returns
    (scale.py:4-20), computed in float32 with numpy for fixtures that bypass torch."""
    Ki = K.astype(np.float32).copy()
    Ki[:, :, :2, :] = Ki[:, :, :2, :] / np.float32(level_div)
    P = E.astype(np.float32).copy()
    P[:, :, :3, :4] = np.matmul(Ki, E[:, :, :3, :4].astype(np.float32))
    return P
