import sys, numpy as np
sys.path.insert(0, '/root/repo')
from mdf_net_b200 import synthetic as syn

def positions(P, v, hyp, H, W):
    # float64 positions (sample coords in source pixel units, align_corners=False shift)
    ref = P[0, 0].astype(np.float64); src = P[0, v].astype(np.float64)
    proj = src @ np.linalg.inv(ref)
    rot, tr = proj[:3, :3], proj[:3, 3]
    y, x = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing='ij')
    r = rot @ np.stack([x.ravel(), y.ravel(), np.ones(H * W)])
    r = r.reshape(3, 1, H, W)
    d = hyp[0].astype(np.float64)  # (D,1,1) or (D,H,W)
    X = r[0] * d + tr[0]; Y = r[1] * d + tr[1]; Z = r[2] * d + tr[2]
    px, py = X / Z, Y / Z
    ix = (px / ((W - 1) / 2)) * (W / 2) - 0.5
    iy = (py / ((H - 1) / 2)) * (H / 2) - 0.5
    return ix, iy

def analyse(h0, w0, N, seed=100, hyp_kind='scene'):
    K, E = syn.camera_rig(1, N, h0, w0, seed=seed)
    for s in range(3):
        H, W = syn.stage_shapes(h0, w0)[s]
        D, G = syn.STAGE_DEPTHS[s], syn.STAGE_GROUPS[s]
        P = syn.projection_matrices(K, E, 2.0 ** (3 - s))
        if s == 0: hyp = syn.uniform_hypos(1, D)
        elif hyp_kind == 'scene': hyp = syn.scene_hypos(1, D, H, W, seed=seed)
        elif hyp_kind == 'wide': hyp = syn.scene_hypos(1, D, H, W, seed=seed, range_mm=(40., 102.))
        else: hyp = syn.pixel_hypos(1, D, H, W, seed=seed)
        print(f'stage {s} {H}x{W} D{D} G{G} hyp={hyp_kind}')
        for v in range(1, N):
            ix, iy = positions(P, v, hyp, H, W)
            dx = np.abs(np.diff(ix, axis=0)); dy = np.abs(np.diff(iy, axis=0))
            sx = np.diff(ix, axis=2); 
            inside = (ix > -1) & (ix < W) & (iy > -1) & (iy < H)
            # warp-level same cell between consecutive planes, warps = 32 px along x (W padded)
            fx, fy = np.floor(ix), np.floor(iy)
            same = (fx[1:] == fx[:-1]) & (fy[1:] == fy[:-1])
            Wp = (W // 32) * 32
            sw = same[:, :, :Wp].reshape(D - 1, H, Wp // 32, 32).all(-1)
            sq = same[:, :, :Wp].reshape(D - 1, H, Wp // 8, 8).all(-1)
            print(f'  v{v}: inside {inside.mean():.3f} |dx/plane| mean {dx.mean():.3f} max {dx.max():.3f} |dy| mean {dy.mean():.3f} '
                  f'xscale {sx.mean():.4f} [{sx.min():.3f},{sx.max():.3f}] lane-same {same.mean():.3f} warp-same {sw.mean():.3f} quarter-same {sq.mean():.3f}')
            # footprint extents for tiles 32xTH over all planes
            for TH, PL in ((2, 8), (2, D), (4, D), (8, D)):
                Hh = (H // TH) * TH
                res = []
                for p0 in range(0, D, PL):
                    a = ix[p0:p0 + PL, :Hh, :Wp].reshape(-1, Hh // TH, TH, Wp // 32, 32)
                    b = iy[p0:p0 + PL, :Hh, :Wp].reshape(-1, Hh // TH, TH, Wp // 32, 32)
                    ex = np.floor(a.max((0, 2, 4))) - np.floor(a.min((0, 2, 4))) + 2
                    ey = np.floor(b.max((0, 2, 4))) - np.floor(b.min((0, 2, 4))) + 2
                    res.append((ex, ey))
                ex = np.stack([r[0] for r in res]); ey = np.stack([r[1] for r in res])
                print(f'     tile 32x{TH} planes {PL}: box w mean {ex.mean():.1f} p99 {np.percentile(ex,99):.0f} max {ex.max():.0f}; h mean {ey.mean():.1f} p99 {np.percentile(ey,99):.0f} max {ey.max():.0f}')

if __name__ == '__main__':
    kind = sys.argv[1] if len(sys.argv) > 1 else 'scene'
    analyse(1152, 1600, 5, hyp_kind=kind)
