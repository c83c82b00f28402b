"""Plain-PyTorch restatement of the differentiable part of the path (float32 or float64), used ONLY by the tests
as the autograd reference of the CUDA backward: VectorAggregate in eval and train mode
(net/unit/homoaggregate.py:8-46 with net/unit/base.py:85-126).  Pinned against gradients of the unmodified
reference by tests/golden/vecagg_grad_*.npz (tests/golden/make_golden.py::gen_vecagg_grad)."""
import torch
import torch.nn.functional as F


def sampling_grid(src_proj, ref_proj, depth_hypos, H, W):
    """Normalised grid (B, D*H, W, 2) the reference hands to grid_sample (no gradient flows through it)."""
    with torch.no_grad():
        B, D = depth_hypos.shape[:2]
        dt, dev = src_proj.dtype, src_proj.device
        proj = src_proj @ torch.linalg.inv(ref_proj)
        rot, trans = proj[:, :3, :3], proj[:, :3, 3]
        ys, xs = torch.meshgrid(torch.arange(H, dtype=dt, device=dev), torch.arange(W, dtype=dt, device=dev), indexing="ij")
        pix = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(H * W, dtype=dt, device=dev)])      # (3, HW)
        ray = rot @ pix                                                                                 # (B, 3, HW)
        depth = depth_hypos.expand(B, D, H, W).reshape(B, 1, D, H * W)
        pts = ray[:, :, None, :] * depth + trans[:, :, None, None]                                      # (B, 3, D, HW)
        u = pts[:, 0] / pts[:, 2] / ((W - 1) / 2) - 1
        v = pts[:, 1] / pts[:, 2] / ((H - 1) / 2) - 1
        return torch.stack([u, v], -1).reshape(B, D * H, W, 2)


def vector_aggregate(features, ref_proj, src_projs, depth_hypos, cw, bn_w, bn_b, bn_mean, bn_var, bn_eps, fc_w, fc_b,
                     groups, training=False, momentum=0.1):
    """Returns (cost volume (B,G,D,H,W), [(batch mean, unbiased batch var) per view] in train mode)."""
    ref, srcs = features[0], features[1:]
    B, C, H, W = ref.shape
    D = depth_hypos.shape[1]
    cpg = C // groups
    ref_unit = F.softmax(ref.reshape(B, groups, cpg, 1, H, W), dim=2)
    num, den, stats = 0.0, 0.0, []
    for fea, sp in zip(srcs, src_projs):
        grid = sampling_grid(sp, ref_proj, depth_hypos, H, W)
        warped = F.grid_sample(fea, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
        unit = F.softmax(warped.reshape(B, groups, cpg, D, H, W), dim=2)
        sim = (unit * ref_unit).sum(2)                                              # (B,G,D,H,W)
        z = (sim * cw.reshape(1, groups, 1, 1, 1)).sum(1, keepdim=True)             # Conv3d(G,1,k=1), no bias
        if training:
            mean, var = z.mean(), z.var(unbiased=False)
            stats.append((mean.detach(), z.var(unbiased=True).detach()))
        else:
            mean, var = torch.as_tensor(bn_mean, dtype=z.dtype, device=z.device), torch.as_tensor(bn_var, dtype=z.dtype, device=z.device)
        h = (z - mean) / torch.sqrt(var + bn_eps) * bn_w + bn_b
        w = torch.sigmoid(torch.relu(h) * fc_w + fc_b)
        num = num + w * sim
        den = den + w
    return num / den, stats
