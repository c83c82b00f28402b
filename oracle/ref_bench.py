"""Timing of the UNMODIFIED reference (oracle/_ref, see ref_install.py) for bench.py's baseline legs -- bench
infrastructure, never a product path.

  cpu_hot_path      the reference's own modules for the hot path (VectorAggregate -> F.softmax -> regress.*; the calls
                    CoreNet.forward makes at core.py:58,64,75-77) on torch CPU with all host threads
                    (`--impl reference`, `cpu_baseline.kind = "reference"`)
  cuda_hot_path     the same modules on the same GPU through ATen / cuDNN (`aten_cuda_baseline`: the bar SURVEY 2b names)
  pipeline          FPN and the three 3-D regularisers timed separately (north_star), and the whole eval forward of
                    config.model (eval.py:23-31) against the same model with this repo's drop-ins injected
"""
from __future__ import annotations

import contextlib
import io
import os
import statistics
import time

import numpy as np


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())     # the reference's constructors print; bench stdout is one JSON line


def _modules(view, device, ref):
    import torch
    mods = []
    for st in view:
        p, G = st["params"], st["G"]
        with _quiet():
            m = ref.homoaggregate.VectorAggregate(G).to(device).eval()
        dw = m.depth_weight
        with torch.no_grad():
            dw[0].conv.weight.copy_(torch.from_numpy(np.asarray(p["cw"], np.float32)).view(1, G, 1, 1, 1))
            dw[0].bn.weight.fill_(float(p["bn_weight"])); dw[0].bn.bias.fill_(float(p["bn_bias"]))
            dw[0].bn.running_mean.fill_(float(p["bn_mean"])); dw[0].bn.running_var.fill_(float(p["bn_var"]))
            dw[1].weight.fill_(float(p["fc_weight"])); dw[1].bias.fill_(float(p["fc_bias"]))
        mods.append(m)
    return mods


def _tensors(view, device):
    import torch
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return [dict(features=[t(f) for f in st["features"]], ref_proj=t(st["ref_proj"]), src_projs=[t(q) for q in st["src_projs"]],
                 hypos=t(st["hypos"]), logits=t(st["logits"]), D=st["D"]) for st in view]


def _hot_path(ref, mods, tens, planes=None):
    """What CoreNet.forward runs for this path, per stage (core.py:58, regular.py:69 / :133, core.py:64, :75-77).
    `planes[s]` limits stage s to its first planes (bounded CPU samples)."""
    import torch
    import torch.nn.functional as F
    out = None
    with torch.no_grad():
        for s, (m, t) in enumerate(zip(mods, tens)):
            d = t["D"] if planes is None else planes[s]
            hyp = t["hypos"][:, :d].contiguous()
            cv = m(t["features"], t["ref_proj"], t["src_projs"], hyp)
            prob = F.softmax(t["logits"][:, :d].contiguous(), dim=1)
            depth = ref.regress.depth_regression(prob, hyp)
            if s == len(mods) - 1:
                conf = ref.regress.confidence_regress(prob)
                conf = F.interpolate(conf.unsqueeze(1), size=None, scale_factor=2, mode="nearest", align_corners=None).squeeze(1)
                out = (depth, conf)
            del cv
    return out


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:  # pragma: no cover
        return os.cpu_count() or 1


def cpu_hot_path(view, steps: int, warmup: int, budget_s: float):
    """Times `steps` steps after `warmup`; every step is the reference's hot path for ONE view restricted to the first
    planes of each stage when a full view would not fit `budget_s` for the whole run (cost is linear in the planes).
    Returns a dict with views/s extrapolated by the fraction of (view, group, plane, pixel) evaluations a step covers."""
    import torch
    from . import ref_install
    with _quiet():
        ref = ref_install.modules()
    threads = host_threads()
    torch.set_num_threads(threads)          # torchrun exports OMP_NUM_THREADS=1
    cpu = torch.device("cpu")
    mods, tens = _modules(view, cpu, ref), _tensors(view, cpu)
    work = [st["G"] * st["H"] * st["W"] for st in view]                      # evaluations per plane and source view
    full = sum(w * st["D"] for w, st in zip(work, view))
    probe_planes = [max(1, st["D"] // 8) for st in view]
    t0 = time.perf_counter()
    _hot_path(ref, mods, tens, probe_planes)
    probe = time.perf_counter() - t0
    probe_frac = sum(w * d for w, d in zip(work, probe_planes)) / full
    est_full = probe / probe_frac
    frac = min(1.0, budget_s / (max(1, steps + warmup) * est_full))
    planes = [max(1, min(st["D"], int(round(st["D"] * frac)))) for st in view]
    if frac >= 1.0:
        planes = [st["D"] for st in view]
    frac = sum(w * d for w, d in zip(work, planes)) / full
    for _ in range(warmup):
        _hot_path(ref, mods, tens, planes)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        _hot_path(ref, mods, tens, planes)
        times.append(time.perf_counter() - t0)
    t = statistics.mean(times)
    batch = int(view[0]["features"][0].shape[0])
    return {"value": batch * frac / t, "unit": "views/s", "cores": threads, "kind": "reference", "host_cpus": os.cpu_count(),
            "torch_threads": torch.get_num_threads(), "s_per_step": t, "fraction_of_a_view_per_step": frac,
            "sample": f"each step = the unmodified reference modules (net/unit/homoaggregate.py VectorAggregate, F.softmax, "
                      f"net/unit/regress.py) on torch {torch.__version__} CPU, {threads} threads, planes {planes} of "
                      f"{[st['D'] for st in view]} per stage of one view = {frac:.3f} of a view, {t:.2f} s/step"}


def cuda_hot_path(view, device, reps: int = 3):
    """The reference's modules for the hot path on `device` (ATen / cuDNN), CUDA events, ms per view (median)."""
    import torch
    from . import ref_install
    with _quiet():
        ref = ref_install.modules()
    mods, tens = _modules(view, device, ref), _tensors(view, device)
    _hot_path(ref, mods, tens)
    torch.cuda.synchronize(device)
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); _hot_path(ref, mods, tens); b.record()
        torch.cuda.synchronize(device)
        ts.append(a.elapsed_time(b))
    ms = statistics.median(ts)
    batch = int(view[0]["features"][0].shape[0])
    peak = torch.cuda.max_memory_allocated(device)
    del mods, tens
    torch.cuda.empty_cache()
    return {"ms_per_view": ms, "value": batch / (ms / 1e3), "unit": "views/s", "reps": reps, "kind": "reference",
            "what": "the unmodified reference modules for the same path (VectorAggregate incl. homo_warping / grid_sample, "
                    "F.softmax, depth_regression, confidence_regress + nearest x2) on the same GPU through ATen / cuDNN, same inputs, "
                    "CUDA events around eager calls; includes the host sync of torch.inverse (base.py:98)",
            "max_memory_allocated_bytes": int(peak)}


def randomise_weights(model, sample_img, seed: int = 11, feature_std: float = 1.5):
    """Seeded stand-in for the missing checkpoints (pth/dtu_29.pth is absent from the checkout).  Default-init features are
    ~1e-4 and give a degenerate cost volume == 0.5, so the BatchNorm statistics are randomised -- and the three
    1x1 output convolutions of the FPN (backbone.py:43-45, bias-free) are rescaled so that the features the cost volume sees
    have a standard deviation of `feature_std`, the magnitude of trained feature maps.  (Without the rescaling the
    randomised BatchNorms stack up to |features| ~ 1e5: every similarity saturates to 0 / 1 and float32 blending noise alone
    flips them.)"""
    import torch
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            with torch.no_grad():
                m.running_var.copy_((torch.rand(m.running_var.shape, generator=g) * 0.04 + 0.002).to(m.running_var.device))
                m.running_mean.copy_((torch.randn(m.running_mean.shape, generator=g) * 0.05).to(m.running_mean.device))
                m.weight.copy_((1.0 + 0.3 * torch.randn(m.weight.shape, generator=g)).to(m.weight.device))
                m.bias.copy_((0.1 * torch.randn(m.bias.shape, generator=g)).to(m.bias.device))
    with torch.no_grad():
        was_training = model.training
        model.eval()
        y4, y3, y2 = model.Backbone(sample_img)
        for conv, y in ((model.Backbone.out4, y4), (model.Backbone.out3, y3), (model.Backbone.out2, y2)):
            conv.weight.mul_(feature_std / float(y.std()))
        model.train(was_training)
    return model


def pipeline(device, h0: int, w0: int, nviews: int, reps: int = 3):
    """FPN and 3-D CNN timed separately (north_star), and the whole eval forward (eval.py:23-31) of config.model next
    to the same model with this repo's units injected (the three config.py lines of INTEGRATION.md + the fused CoreNet)."""
    import torch
    import mdf_net_b200 as mdf
    from mdf_net_b200 import synthetic as syn
    from . import ref_install

    with _quiet():
        model = ref_install.config_model()
        ref = ref_install.modules()
    model = model.to(device).eval()
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True          # what config.py:32 sets for eval
    K, E = syn.camera_rig(1, nviews, h0, w0, seed=5)
    rng = np.random.default_rng(5)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    imgs = t(rng.random((1, nviews, 3, h0, w0), dtype=np.float32))
    args = (imgs, t(E), t(K), t(np.array([[425.0, 935.0]], np.float32)))
    randomise_weights(model, imgs[:, 0])

    def med(fn, n=reps, warm=1):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(device)
        ts = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize(device)
            ts.append(a.elapsed_time(b))
        return statistics.median(ts)

    out = {"not_in_value": True, "workload": f"{w0}x{h0} N={nviews}, batch 1, seeded weights with randomised BatchNorm statistics and "
                                           f"unit-scale features (pth/dtu_29.pth is absent from the checkout), fp32 (TF32 off), cudnn.benchmark on"}
    try:
        with torch.no_grad():
            views = torch.unbind(imgs, 1)
            out["fpn_ms"] = med(lambda: [model.Backbone(v) for v in views])                       # core.py:42, all N views
            reg = []
            for s, (h, w) in enumerate(syn.stage_shapes(h0, w0)):
                cv = torch.rand((1, syn.STAGE_GROUPS[s], syn.STAGE_DEPTHS[s], h, w), device=device)
                reg.append(med(lambda: model.Regular[s](cv)))                                     # core.py:61
                del cv
            out["regulariser_ms"] = reg
            out["refine_ms"] = med(lambda: model.Refine(torch.rand((1, h0 // 2, w0 // 2), device=device) * 500 + 425, args[3]))
            ref_ms = med(lambda: model(*args))
            a = model(*args)
            # the same model, this repo's units injected (state dict shared: the drop-ins keep the reference's keys)
            ours = mdf.CoreNet(model.Backbone, torch.nn.ModuleList([mdf.HyposByFit(h.ndepths, h.curve_calss, float(h.prob_thresh))
                                                                  for h in model.Depth_hypos]),
                               model.scale, torch.nn.ModuleList([mdf.VectorAggregate(g_) for g_ in syn.STAGE_GROUPS]), model.Regular,
                               [mdf.depth_regression, mdf.confidence_regress], model.Refine).to(device).eval()
            ours.Homoaggre.load_state_dict(model.Homoaggre.state_dict(), strict=True)
            ours_ms = med(lambda: ours(*args))
            b = ours(*args)
            err = (a["depth"] - b["depth"]).abs()
            out.update({
                "whole_view_reference_ms": ref_ms, "whole_view_reference_views_per_s": 1e3 / ref_ms,
                "whole_view_dropin_ms": ours_ms, "whole_view_dropin_views_per_s": 1e3 / ours_ms,
                "speedup": ref_ms / ours_ms,
                "depth_within_0.5mm": float((err < 0.5).float().mean()),
                "confidence_mask_agreement_0.8": float(((a["confidence"] > 0.8) == (b["confidence"] > 0.8)).float().mean()),
                "note": "whole eval forward of config.model (eval.py:23-31: FPN x N, 3 x (hypotheses, cost volume, 3-D CNN, regression), "
                        "refine, confidence) vs the same weights with mdf_net_b200's VectorAggregate / HyposByFit / regress / fused "
                        "CoreNet tails injected; CUDA events, eager, median"})
            # the reference out of the box: config.py never touches allow_tf32, so its cuDNN convolutions take PyTorch's default
            # (TF32 tensor cores allowed).  The units of this package compute in fp32 either way; only the clock is reported here.
            torch.backends.cudnn.allow_tf32 = True
            out["tf32_convolutions"] = {"whole_view_reference_ms": med(lambda: model(*args)), "whole_view_dropin_ms": med(lambda: ours(*args)),
                                        "fpn_ms": med(lambda: [model.Backbone(v) for v in views]),
                                        "note": "torch.backends.cudnn.allow_tf32 = True (PyTorch's default, which config.py leaves alone): "
                                                "the FPN and the 3-D CNN run on TF32 tensor cores in both arms"}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    return out
