import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/tmp/an')
from mdf_net_b200 import synthetic as syn
from geom import positions

def leftover(ix, iy, H, W, TH, BW, BH, planes=None):
    D = ix.shape[0]
    inside = (ix > -1) & (ix < W) & (iy > -1) & (iy < H)
    x0 = np.floor(np.clip(ix, -1, W - 1)); y0 = np.floor(np.clip(iy, -1, H - 1))
    Wp, Hp = (W // 32) * 32, (H // TH) * TH
    def tiles(a): return a[:, :Hp, :Wp].reshape(D, Hp // TH, TH, Wp // 32, 32)
    X, Y, I = tiles(x0), tiles(y0), tiles(inside)
    big = 1e9
    ox = np.where(I, X, big).min((0, 2, 4), keepdims=True); oy = np.where(I, Y, big).min((0, 2, 4), keepdims=True)
    fit = ((X - ox) < BW - 1) & ((Y - oy) < BH - 1)
    left = I & ~fit
    frac = left.sum() / I.sum()
    warp = left.any(-1).sum() / max(1, I.any(-1).sum())
    tile = left.any((0, 2, 4)).mean()
    return frac, warp, tile

def run(kind, cfgs):
    h0, w0, N, seed = 1152, 1600, 5, 100
    K, E = syn.camera_rig(1, N, h0, w0, seed=seed)
    for s, lst in cfgs.items():
        H, W = syn.stage_shapes(h0, w0)[s]; D = syn.STAGE_DEPTHS[s]
        P = syn.projection_matrices(K, E, 2.0 ** (3 - s))
        hyp = syn.uniform_hypos(1, D) if s == 0 else (syn.scene_hypos(1, D, H, W, seed=seed) if kind == 'scene' else
              syn.scene_hypos(1, D, H, W, seed=seed, range_mm=(40., 102.)) if kind == 'wide' else syn.pixel_hypos(1, D, H, W, seed=seed))
        pos = [positions(P, v, hyp, H, W) for v in range(1, N)]
        for (TH, BW, BH, SL) in lst:
            out = []
            for ix, iy in pos:
                fr = []
                for p0 in range(0, D, SL):
                    fr.append(leftover(ix[p0:p0+SL], iy[p0:p0+SL], H, W, TH, BW, BH))
                out.append(np.mean(fr, 0))
            out = np.array(out)
            print(f'stage {s} {kind} tile 32x{TH} slab {SL} box {BW}x{BH}: leftover sample-views per view ' + ' '.join(f'{x:.4f}' for x in out[:, 0]) +
                  ' | warp-samples ' + ' '.join(f'{x:.4f}' for x in out[:, 1]) + ' | tiles ' + ' '.join(f'{x:.3f}' for x in out[:, 2]))

if __name__ == '__main__':
    kind = sys.argv[1]
    run(kind, {0: [(2, 40, 6, 4), (2, 48, 6, 4), (2, 48, 6, 8), (4, 40, 8, 4)],
               1: [(4, 44, 8, 24), (4, 48, 8, 24), (4, 40, 8, 12), (2, 48, 6, 24), (4, 56, 10, 24), (8, 40, 12, 24)],
               2: [(8, 56, 12, 8), (8, 48, 12, 8), (4, 56, 8, 8), (4, 64, 10, 8), (8, 64, 14, 8)]})
