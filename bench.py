#!/usr/bin/env python
"""bench.py -- throughput of MDF-Net's plane-sweep cost-volume path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W     # the reference algorithm on the host CPU cores

One "step" = the hot path for ONE reference view of BASELINE.json configs[1] (DTU eval 1600x1152,
N=5 views, batch 1): for each of the 3 cost-volume stages (1/8, 1/4, 1/2 resolution;
C=64/32/16, D=48/24/8, G=32/16/8) the fused warp + aggregate kernel, then the fused softmax +
depth-regression (+ confidence on the last stage) kernel on that stage's regulariser logits.
The 2-D feature pyramid and the 3-D regulariser are NOT part of the path (they stay on
PyTorch/cuDNN, north_star): features and logits are synthetic, seeded, of the real shapes.

Prints ONE JSON line (rank 0).  `value` = views/s with inputs resident in HBM (all ranks' views /
max-over-ranks device time); `e2e` = views/s through the public Python plugin API with pinned HOST
inputs (H2D of features/projections/hypotheses/logits and D2H of depth + confidence inside the
timed region); `roofline` = algorithmic HBM bytes of the fused cost-volume launches / their CUDA-event
time vs MEASURED_PEAKS.json; `cpu_baseline` = the CPU oracle (port of the reference algorithm) on
this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

from mdf_net_b200 import synthetic as syn

WORKLOADS = {
    # name: (H0, W0, N views, batch)
    "dtu_1600x1152_n5": (1152, 1600, 5, 1),      # BASELINE.json configs[1] / [2]  (metric config)
    "dtu_1600x1184_n5": (1184, 1600, 5, 1),      # the crop the shipped loader really uses (dtueval.py:34)
    "dtu_640x512_n3": (512, 640, 3, 1),          # configs[0]
    "tanks_1920x1056_n7": (1056, 1920, 7, 1),    # configs[3]
    "tanks_1920x1056_n11": (1056, 1920, 11, 1),  # configs[3] with the reference's own default N (config.py:119)
}
LAUNCHES_PER_STEP = 3 * (3 + 1)   # per stage: setup + prep + staged cost-volume kernel, + the fused head kernel
FALLBACK_HBM_GBS = 6650.0
E2E_PASSES = 5
CURVES, PROB_THRESH = (None, "gauss1", "laplace"), (0.0, 0.95, 1e-5)      # config.py:200-201
# ncu --set full summaries of the three cost_volume_staged_kernel launches of one step, per workload (newest first)
NCU_SUMMARIES = {"dtu_1600x1152_n5": ["profiles/r02_final_staged_ncu_full_summary.txt", "profiles/r01_final_staged_ncu_full_summary.txt"]}


def ncu_dram_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of the hot kernel's launches of one step, parsed from the committed
    ncu summary of this workload (None when there is none).  It sits below the algorithmic bytes: the layout pass leaves
    the difference maps in L2 and part of the volume is still dirty in L2 when the kernel ends."""
    for rel in NCU_SUMMARIES.get(workload, []):
        try:
            total, found = 0.0, 0
            for line in open(os.path.join(ROOT, rel)):
                if line.startswith(("dram__bytes_read.sum [Mbyte]:", "dram__bytes_write.sum [Mbyte]:")):
                    total += sum(float(x) for x in line.split(":", 1)[1].split("|")) * 1e6
                    found += 1
            if found == 2:
                return {"bytes": total, "source": rel}
        except OSError:
            continue
    return None


# ----------------------------------------------------------------------------------------- workload
def make_view(h0, w0, nviews, batch, seed, chain=False):
    """Host (numpy) inputs of one reference view: per stage features, projections, hypotheses,
    depth_weight parameters and regulariser logits.  chain=True: logits that make the coarse-to-fine chain behave
    like a scene when the hypotheses of stages 1-2 are produced on the device (the end-to-end leg)."""
    K, E = syn.camera_rig(batch, nviews, h0, w0, seed=seed)
    stages = []
    for s in range(3):
        H, W = syn.stage_shapes(h0, w0)[s]
        C, D, G = syn.STAGE_CHANNELS[s], syn.STAGE_DEPTHS[s], syn.STAGE_GROUPS[s]
        P = syn.projection_matrices(K, E, 2.0 ** (3 - s))
        stages.append(dict(
            H=H, W=W, C=C, D=D, G=G,
            features=syn.smooth_features(batch, nviews, C, H, W, seed=seed + 10 + s),
            ref_proj=P[:, 0].copy(), src_projs=[P[:, v].copy() for v in range(1, nviews)],
            hypos=syn.uniform_hypos(batch, D) if s == 0 else syn.scene_hypos(batch, D, H, W, seed=seed),
            params=syn.depth_weight_params(G, seed=seed + 30 + s),
            logits=syn.scene_logits(batch, s, H, W, seed=seed) if chain else syn.regulariser_logits(batch, D, H, W, seed=seed + 40 + s)))
    return stages


def algorithmic_bytes(h0, w0, nviews, batch):
    """per stage 4*B*[N*C*H*W + D*(1|H*W) + G*D*H*W] (cost volume) and the head's
    4*B*[D*H*W logits + D*(1|H*W) hypotheses + D*H*W prob + H*W depth] (+ H0*W0 confidence, last stage)."""
    cv, head = [], []
    for s in range(3):
        H, W = syn.stage_shapes(h0, w0)[s]
        C, D, G = syn.STAGE_CHANNELS[s], syn.STAGE_DEPTHS[s], syn.STAGE_GROUPS[s]
        hyp = D * (1 if s == 0 else H * W)
        cv.append(4 * batch * (nviews * C * H * W + hyp + G * D * H * W))
        head.append(4 * batch * (2 * D * H * W + hyp + H * W + (4 * H * W if s == 2 else 0)))
    return cv, head


def onchip_bound(h0, w0, nviews, batch, sm_mhz, hot_ms):
    """The gather core's own ceiling: every (source view, group, depth, pixel) evaluation reads 4 taps
    x 4 B from shared memory; 148 SMs x 128 B/clk.  Reported next to the HBM roofline because it, not HBM, bounds
    the kernel (ncu: l1tex data pipe 69 / 57 / 45 %, DRAM 10-16 %)."""
    evals = sum((nviews - 1) * g * d * h * w * batch for (h, w), d, g in
                zip(syn.stage_shapes(h0, w0), syn.STAGE_DEPTHS, syn.STAGE_GROUPS))
    floor_ms = evals * 16.0 / (148 * 128 * sm_mhz * 1e6) * 1e3
    return {"resource": "shared-memory gather bandwidth", "group_evaluations_per_step": evals, "bytes_per_evaluation": 16,
            "floor_ms": floor_ms, "frac": floor_ms / hot_ms if hot_ms > 0 else None}


def make_config(workload):
    h0, w0, nviews, batch = WORKLOADS[workload]
    stages = [f"{h}x{w} C{c} D{d} G{g}" for (h, w), c, d, g in
              zip(syn.stage_shapes(h0, w0), syn.STAGE_CHANNELS, syn.STAGE_DEPTHS, syn.STAGE_GROUPS)]
    return {"workload": workload, "views": nviews, "batch": batch, "stages": stages,
            "step": "3 x (fused cost volume + fused softmax/regress head) for one reference view",
            "sharding": "independent reference views per rank, no data-path collective",
            "l2": "no flush: one step touches 0.83 GB (> 126 MB L2) and two input sets alternate"}


# ------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    """Polls NVML (SM clock, power, throttle reasons) from a thread every ~20 ms while the timed regions run;
    falls back to an `nvidia-smi -lms` subprocess if NVML is not importable."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4,
               "hw_power_brake_slowdown": 0x80}

    def __init__(self, index: int):
        self.index, self.thread, self.stop_flag = index, None, False
        self.sm, self.power, self.reasons, self.max_sm = [], [], set(), None

    def _resolve_handle(self, nv):
        # CUDA_VISIBLE_DEVICES may remap indices: match by PCI bus id of the torch device when possible
        try:
            import torch
            bus = torch.cuda.get_device_properties(self.index).pci_bus_id
            dom = torch.cuda.get_device_properties(self.index).pci_domain_id
            dev = torch.cuda.get_device_properties(self.index).pci_device_id
            return nv.nvmlDeviceGetHandleByPciBusId(f"{dom:08x}:{bus:02x}:{dev:02x}.0".encode())
        except Exception:
            return nv.nvmlDeviceGetHandleByIndex(self.index)

    def _run(self):
        import pynvml as nv
        h = self._resolve_handle(nv)
        try:
            self.max_sm = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        except Exception:
            pass
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None

    def stop(self) -> dict:
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable"]}
        self.stop_flag = True
        self.thread.join(timeout=2)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_sm,
                "power_w_max": max(self.power) if self.power else None, "samples": len(self.sm),
                "reasons": sorted(self.reasons), "source": "NVML polled every 20 ms across the timed regions (eager pass, graph replays, e2e)"}


# ------------------------------------------------------------------------------------- CPU baseline
def cpu_sample(view, frac_rows: float):
    """The reference algorithm (CPU oracle, OpenMP over the host cores) on the first `frac_rows` of the
    rows of every stage of one view: cost volume, softmax, depth regression, confidence.  Returns seconds."""
    from oracle import c_oracle as co
    t0 = time.perf_counter()
    for s, st in enumerate(view):
        y1 = max(1, int(round(st["H"] * frac_rows)))
        co.vector_aggregate(st["features"], st["hypos"], st["params"], st["G"], ref_proj=st["ref_proj"],
                            src_projs=st["src_projs"], rows=(0, y1))
        logits = np.ascontiguousarray(st["logits"][:, :, :y1])
        hyp = st["hypos"] if s == 0 else np.ascontiguousarray(st["hypos"][:, :, :y1])
        prob = co.softmax_depth(logits)
        co.depth_regression(prob, hyp)
        if s == 2:
            co.confidence_regress(prob, upsample=2)
    return time.perf_counter() - t0


def port_sample(view, budget_s: float, batch: int):
    """The C restatement (oracle/mdf_oracle.c, OpenMP over the host cores) on a bounded sample: kept next to the reference's
    own number as a second opinion (it is ~10x faster than the reference's torch-CPU path on the same cores)."""
    from oracle import c_oracle as co
    co.build()
    co.set_num_threads(co.host_threads())
    probe = cpu_sample(view, 1.0 / 16.0)                     # calibrate: 1/16 of the rows
    frac = min(1.0, max(1.0 / 16.0, budget_s / (probe * 16.0)))
    t = cpu_sample(view, frac)
    return {"value": batch * frac / t, "unit": "views/s", "cores": co.num_threads(), "kind": "port",
            "sample": f"first {frac:.3f} of the rows of all 3 stages of one view (cost volume + head), {t:.2f} s, "
                      f"C oracle oracle/mdf_oracle.c with OpenMP"}


def cpu_baseline(view, budget_s: float, batch: int):
    """The reference's own PyTorch modules for the path on this box's host cores (oracle/_ref, unmodified; kind
    "reference"); the C port only where the reference copy did not travel (kind "port")."""
    from oracle import ref_install
    if ref_install.available():
        from oracle import ref_bench
        out = ref_bench.cpu_hot_path(view, steps=1, warmup=0, budget_s=budget_s)
        try:
            out["port_c_openmp"] = port_sample(view, min(budget_s, 4.0), batch)
        except Exception as e:  # pragma: no cover
            out["port_c_openmp"] = {"error": f"{type(e).__name__}: {e}"}
        return out
    return port_sample(view, budget_s, batch)


def run_reference(args, workload, out):
    """--impl reference: the reference's own implementation of the path on the host cores -- the unmodified
    VectorAggregate / F.softmax / regress.* of oracle/_ref on torch CPU with every host thread (kind "reference").  Each step is
    a bounded sample of one view sized so that the whole run stays within ~2.5 minutes.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import ref_install
    h0, w0, nviews, batch = WORKLOADS[workload]
    view = make_view(h0, w0, nviews, batch, seed=1)
    if ref_install.available():
        from oracle import ref_bench
        base = ref_bench.cpu_hot_path(view, steps=args.steps, warmup=args.warmup, budget_s=150.0)
        value, t, frac = base["value"], base["s_per_step"], base["fraction_of_a_view_per_step"]
    else:                                    # the reference copy did not travel: the C port, said so in `kind`
        from oracle import c_oracle as co
        co.build()
        co.set_num_threads(co.host_threads())      # torchrun exports OMP_NUM_THREADS=1
        probe = cpu_sample(view, 1.0 / 16.0)
        total = max(1, args.steps + args.warmup)
        frac = min(1.0, max(1.0 / 32.0, (150.0 / total) / (probe * 16.0)))     # whole run within ~2.5 minutes
        for _ in range(args.warmup):
            cpu_sample(view, frac)
        times = [cpu_sample(view, frac) for _ in range(args.steps)]
        t = sum(times) / len(times)
        value = batch * frac / t
        base = {"value": value, "unit": "views/s", "cores": co.num_threads(), "kind": "port",
                "sample": f"each step = first {frac:.3f} of the rows of all 3 stages of one view, {t:.2f} s/step (C port: oracle/_ref missing)"}
    out.emit(json.dumps({
        "impl": "reference", "metric": "DTU views/s (plane-sweep cost-volume path)", "value": value, "unit": "views/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / frac,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(workload),
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def time_regulariser_tail(dev, h0, w0, batch, host_view):
    """SURVEY 8f rows 1-2, reported next to the headline (not part of `value`): the fused tail of the regulariser
    (mdf_prob_head_fwd: prob conv + softmax + depth regression + confidence / curve fit, one launch per stage) on
    synthetic post-ReLU feature volumes of the real shapes (c0 = 16 / 8 / 8, config.py:208-212), and what the same three
    convolutions cost through cuDNN on this GPU (the reference's path for that layer).  CUDA-graph replays, us."""
    import torch
    import torch.nn.functional as F
    from mdf_net_b200 import ops

    def graph_us(fn, reps=10, n=5):
        fn(); torch.cuda.synchronize()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        g, keep = torch.cuda.CUDAGraph(), []
        with torch.cuda.graph(g):
            for _ in range(reps):
                keep.append(fn())
        g.replay(); torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / reps * 1e3)
        return statistics.median(ts)

    fused, cudnn, nbytes = [], [], []
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for s, st in enumerate(host_view):
            c0 = (16, 8, 8)[s]
            gen = torch.Generator(device=dev).manual_seed(50 + s)
            x = torch.randn((batch, c0, st["D"], st["H"], st["W"]), device=dev, generator=gen).relu_()
            w = torch.randn((1, c0, 3, 3, 3), device=dev, generator=gen) * 0.35
            hyp = torch.from_numpy(np.ascontiguousarray(st["hypos"])).to(dev)
            curve = ("gauss1", "laplace", "")[s]
            fused.append(graph_us(lambda: ops.prob_head(x, w, hyp, curve, False, False, s == 2)))
            cudnn.append(graph_us(lambda: F.conv3d(x, w, padding=1), reps=3, n=3))
            nbytes.append(x.numel() * 4)
            del x
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    return {"what": "prob conv + softmax + depth regression + confidence / curve fit, one launch per stage (mdf_prob_head_fwd)",
            "fused_us": fused, "fused_us_per_view": sum(fused), "input_bytes": nbytes,
            "GBps": [b / 1e9 / (t / 1e6) for b, t in zip(nbytes, fused)],
            "cudnn_conv3d_alone_us": cudnn, "not_in_value": True}


def time_other_rows(dev):
    """Other rows of the scope table, reported next to the headline (not part of `value`): train-mode forward + backward
    of the fused op at BASELINE.json configs[4] (768x576, N=5, batch 8), `homo_warping` at configs[1] stage 0 and the
    geometric-consistency filter at 1600x1200 with 10 source views.  CUDA events around eager calls, median of 5, ms."""
    import torch
    import mdf_net_b200 as mdf
    from mdf_net_b200 import ops

    def med(fn, n=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts)

    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = {"not_in_value": True}
    h0, w0, nviews, batch = 576, 768, 5, 8
    K, E = syn.camera_rig(batch, nviews, h0, w0, seed=3)
    train, steady = [], []
    for s in range(3):
        H, W = syn.stage_shapes(h0, w0)[s]
        C, D, G = syn.STAGE_CHANNELS[s], syn.STAGE_DEPTHS[s], syn.STAGE_GROUPS[s]
        P = syn.projection_matrices(K, E, 2.0 ** (3 - s))
        feats = [cu(f).requires_grad_(True) for f in syn.smooth_features(batch, nviews, C, H, W, seed=60 + s)]
        hyp = cu(syn.uniform_hypos(batch, D) if s == 0 else syn.pixel_hypos(batch, D, H, W, seed=61 + s))
        rp, sps = cu(P[:, 0]), [cu(P[:, v]) for v in range(1, nviews)]
        m = mdf.VectorAggregate(G).to(dev).train()
        go = torch.randn((batch, G, D, H, W), device=dev)
        train.append(med(lambda: m(feats, rp, sps, hyp).backward(go), n=3, warm=1))
        # steady state: four iterations follow each other (the CPU runs ahead of the GPU, as in a training loop); the single-call
        # figure above starts every measurement on an idle GPU and so includes ~0.2 ms of Python / autograd before the first launch
        steady.append(med(lambda: [m(feats, rp, sps, hyp).backward(go) for _ in range(4)], n=3, warm=1) / 4.0)
        del feats, go, m
    out["train_fwd_bwd_ms_768x576_n5_b8"] = train
    out["train_fwd_bwd_steady_state_ms_768x576_n5_b8"] = steady
    K, E = syn.camera_rig(1, 5, 1152, 1600, seed=1)
    H, W = syn.stage_shapes(1152, 1600)[0]
    P = syn.projection_matrices(K, E, 8.0)
    f = cu(syn.smooth_features(1, 2, 64, H, W, seed=10)[1])
    hyp = cu(syn.uniform_hypos(1, 48))
    sp, rp = cu(P[:, 1]), cu(P[:, 0])                     # resident, like the features: the row times the op, not two uploads
    t = med(lambda: [ops.homo_warp(f, sp, rp, hyp) for _ in range(10)]) / 10.0      # back to back: the CPU runs ahead, as in a pipeline
    out["homo_warp_stage0"] = {"ms": t, "output_GBps": 64 * 48 * H * W * 4 / 1e9 / (t / 1e3), "calls_back_to_back": 10}
    # the FPN hand-off: one view's three cost volumes from prepared maps (setup + hot kernel, no layout pass),
    # next to the NCHW drop-in entry on features of the same content; and the library's 1x1 output convolutions (N views)
    try:
        K, E = syn.camera_rig(1, 5, 1152, 1600, seed=1)
        nchw, prepped, conv1x1 = [], [], []
        for s in range(3):
            H, W = syn.stage_shapes(1152, 1600)[s]
            C, D, G = syn.STAGE_CHANNELS[s], syn.STAGE_DEPTHS[s], syn.STAGE_GROUPS[s]
            P = syn.projection_matrices(K, E, 2.0 ** (3 - s))
            gen = torch.Generator(device=dev).manual_seed(70 + s)
            xs = [torch.nn.functional.avg_pool2d(torch.randn((1, 64, H, W), device=dev, generator=gen), 3, 1, 1) * 3.0 for _ in range(5)]
            wt = torch.randn((C, 64), device=dev, generator=gen) * 0.08
            hyp = cu(syn.uniform_hypos(1, D) if s == 0 else syn.scene_hypos(1, D, H, W, seed=1))
            m = mdf.VectorAggregate(G).to(dev).eval()
            cwt = m.depth_weight[0].conv.weight
            rp, sps = cu(P[:, 0]), [cu(P[:, v]) for v in range(1, 5)]
            feats = [torch.nn.functional.conv2d(x, wt.view(C, 64, 1, 1)) for x in xs]
            with torch.no_grad():
                conv1x1.append(med(lambda: [ops.fpn_out_prepped(x, wt, G, cwt, i == 0) for i, x in enumerate(xs)]))
                q4, cq4 = ops.fpn_out_prepped(xs[0], wt, G, cwt, True)
                s4 = torch.stack([ops.fpn_out_prepped(x, wt, G, cwt, False)[0] for x in xs[1:]], 0)
                nchw.append(med(lambda: m(feats, rp, sps, hyp)))
                prepped.append(med(lambda: m(mdf.PreppedFeatures(q4, cq4, s4), rp, sps, hyp)))
            del xs, feats, s4, q4, cq4
        out["fpn_handoff_1600x1152_n5"] = {"cost_volume_nchw_entry_ms": nchw, "cost_volume_prepped_entry_ms": prepped,
                                           "library_1x1_output_convs_5_views_ms": conv1x1,
                                           "note": "eager calls, CUDA events, median of 5; the prepared entry runs setup + hot kernel (no prep_kernel)"}
    except Exception as e:  # pragma: no cover
        out["fpn_handoff_1600x1152_n5"] = {"error": f"{type(e).__name__}: {e}"}
    S, Hf, Wf = 10, 1200, 1600
    Kf, Ef = syn.camera_rig(1, S + 1, Hf, Wf, seed=321)
    gen = torch.Generator(device=dev).manual_seed(1)
    depths = [700.0 + 20.0 * torch.rand((Hf, Wf), device=dev, generator=gen) for _ in range(S + 1)]
    conf = torch.rand((Hf, Wf), device=dev, generator=gen)
    Kt, Et = cu(Kf[0]), cu(Ef[0])
    out["geo_filter_1600x1200_10src_ms"] = med(lambda: ops.geo_filter(depths[0], Kt[0], Et[0], depths[1:], Kt[1:], Et[1:], conf,
                                                                        0.8, 3, 4.0, 1300.0))
    return out


def pin_to_gpu_cpus(index: int):
    """Bind this process (and the pinned arenas it allocates afterwards: first touch) to the CPUs NVML reports as local to
    the GPU.  Returns the number of CPUs bound to, or None when NVML / the affinity call is unavailable."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        words = nv.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def time_other_workloads(dev, names, peak):
    """The other configurations of BASELINE.json (and the reference's own defaults), one view each: eager hot path, CUDA
    events; views/s and the dominant kernel's share of the HBM roofline.  Not part of `value`."""
    import torch
    from mdf_net_b200 import ops
    rows = {}
    for name in names:
        h0, w0, nviews, batch = WORKLOADS[name]
        view = make_view(h0, w0, nviews, batch, seed=3)
        cvb, _ = algorithmic_bytes(h0, w0, nviews, batch)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        f32 = lambda v: t(np.asarray(v, np.float32).reshape(-1))
        dv = [dict(G=st["G"], features=[t(f) for f in st["features"]], ref_proj=t(st["ref_proj"]), src_projs=[t(q) for q in st["src_projs"]],
                   hypos=t(st["hypos"]), logits=t(st["logits"]),
                   w=[f32(st["params"][k]) for k in ("cw", "bn_weight", "bn_bias", "bn_mean", "bn_var", "fc_weight", "fc_bias")],
                   eps=float(st["params"]["bn_eps"])) for st in view]

        def step(kev=None):
            for s, st in enumerate(dv):
                if kev is not None:
                    ops.time_next_hot_kernel(*kev[s])
                cv = ops.cost_volume(st["features"], st["ref_proj"], st["src_projs"], st["hypos"], *st["w"][:5], st["eps"], st["w"][5], st["w"][6], st["G"], 0)
                ops.softmax_regress(st["logits"], st["hypos"], True, s == 2, 4, 1, 2, 2)
                del cv
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        ms, hot = [], []
        for _ in range(5):
            kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(kev); b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
            hot.append([e0.elapsed_time(e1) for e0, e1 in kev])
        m = statistics.median(ms)
        hk = [statistics.median(x[s] for x in hot) for s in range(3)]
        rows[name] = {"views_per_s": batch / (m / 1e3), "ms_per_view": m, "hot_kernel_us": [1e3 * x for x in hk],
                      "algorithmic_bytes": sum(cvb), "hot_kernel_GBps": sum(cvb) / 1e9 / (sum(hk) / 1e3),
                      "roofline_frac": sum(cvb) / 1e9 / (sum(hk) / 1e3) / peak, "launch_mode": "eager"}
        del dv
        torch.cuda.empty_cache()
    rows["not_in_value"] = True
    return rows


# ----------------------------------------------------------------------------------------- GPU arm
def run_b200(args, workload, out):
    import torch
    import torch.distributed as dist

    import mdf_net_b200 as mdf
    from mdf_net_b200 import _cabi, ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this path has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    bound_cpus = pin_to_gpu_cpus(local)      # before any pinned allocation: the arenas land on the GPU's NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    h0, w0, nviews, batch = WORKLOADS[workload]
    # every rank owns its own reference views (shard by view, no data-path collective); two distinct
    # input sets alternate so consecutive steps never see the same buffers
    host_views = [make_view(h0, w0, nviews, batch, seed=1000 * rank + 17 * i + 1) for i in range(2)]
    cv_bytes, head_bytes = algorithmic_bytes(h0, w0, nviews, batch)

    def to_dev(a, pin=False):
        t = torch.from_numpy(np.ascontiguousarray(a))
        return t.pin_memory() if pin else t.to(dev)

    def convert(view, pin):
        out = []
        for st in view:
            p = st["params"]
            f32 = lambda v: np.asarray(v, np.float32).reshape(-1)
            out.append(dict(
                G=st["G"], features=[to_dev(f, pin) for f in st["features"]], ref_proj=to_dev(st["ref_proj"], pin),
                src_projs=[to_dev(q, pin) for q in st["src_projs"]], hypos=to_dev(st["hypos"], pin),
                logits=to_dev(st["logits"], pin),
                # depth_weight parameters are model weights: resident on the device in both modes
                w=[to_dev(f32(p[k])) for k in ("cw", "bn_weight", "bn_bias", "bn_mean", "bn_var", "fc_weight", "fc_bias")],
                eps=float(p["bn_eps"])))
        return out

    dev_views = [convert(v, pin=False) for v in host_views]
    # end-to-end leg: the hypotheses of stages 1-2 are NOT shipped; the host sends depth_range and the stage loop forms them
    # on the device (HyposByFit kernels, core.py:55 order), so its logits are the chain-consistent ones
    chain_views = [make_view(h0, w0, nviews, batch, seed=1000 * rank + 17 * i + 1, chain=True) for i in range(2)] if not args.no_e2e else None
    pin_views = [convert(v, pin=True) for v in chain_views] if not args.no_e2e else None
    depth_range_host = torch.tensor([list(syn.DTU_DEPTH_RANGE)] * batch, dtype=torch.float32)

    def hot_path(view, events=None, kernel_events=None):
        """3 x (fused cost volume -> fused head).  Returns the per-stage depth maps and the confidence."""
        depths, conf = [], None
        for s, st in enumerate(view):
            if kernel_events is not None:      # events around the hot kernel alone, recorded inside the library
                ops.time_next_hot_kernel(kernel_events[s][0], kernel_events[s][1])
            if events is not None:
                events[s][0].record()
            cv = ops.cost_volume(st["features"], st["ref_proj"], st["src_projs"], st["hypos"], *st["w"][:5], st["eps"],
                                 st["w"][5], st["w"][6], st["G"], 0)
            if events is not None:
                events[s][1].record()
            # the 3-D regulariser (cuDNN, out of scope) would turn `cv` into logits here
            prob, depth, c = ops.softmax_regress(st["logits"], st["hypos"], True, s == 2, 4, 1, 2, 2)
            depths.append(depth)
            if s == 2:
                conf = c
            del cv, prob
        return depths, conf

    def hot_path_chain(view, depth_range):
        """The stage loop as CoreNet.forward runs it (core.py:45-65) with everything of this path on the device: uniform
        hypotheses from depth_range, then per stage fused cost volume -> fused head that also fits the next stage's curve ->
        next hypotheses (x2 upsampling, range, clamps)."""
        depths, conf, depth, fitted, hyp = [], None, None, None, None
        for s, st in enumerate(view):
            D = syn.STAGE_DEPTHS[s]
            if s == 0:
                hyp = stage0_hypos(depth_range)
            else:
                hyp = ops.hypos_generate(depth, fitted, depth_range, CURVES[s], PROB_THRESH[s], D, True)
            cv = ops.cost_volume(st["features"], st["ref_proj"], st["src_projs"], hyp, *st["w"][:5], st["eps"],
                                 st["w"][5], st["w"][6], st["G"], 0)
            if s < 2:
                _, depth, _, fitted = ops.softmax_regress_fit(st["logits"], hyp, CURVES[s + 1], False, False, 4, 1, 2, 2)
            else:
                _, depth, conf = ops.softmax_regress(st["logits"], hyp, False, True, 4, 1, 2, 2)
            depths.append(depth)
            del cv
        return depths, conf

    hyp0_unit = mdf.HyposByFit(syn.STAGE_DEPTHS[0], None, 0.0)
    def stage0_hypos(depth_range):
        return hyp0_unit(None, depth_range, None, None)

    copy_stream = torch.cuda.Stream(device=dev)
    copy_events = []          # (start, stop) around every arena transfer of the timed e2e passes

    def pack(pview):
        """All host tensors of one view in ONE pinned arena (256-byte aligned slots): the step's inputs then cross PCIe
        as a single transfer instead of 36 (15 of them 64-byte matrices whose copies cost latency, not bandwidth)."""
        slots, off = [], 0
        def add(t):
            nonlocal off
            slots.append((off, t))
            off += (t.numel() + 63) // 64 * 64
            return len(slots) - 1
        layout = [dict(G=st["G"], features=[add(f) for f in st["features"]], ref_proj=add(st["ref_proj"]),
                       src_projs=[add(q) for q in st["src_projs"]], logits=add(st["logits"]),
                       w=st["w"], eps=st["eps"]) for st in pview]
        layout.append(add(depth_range_host))
        arena = torch.empty(off, dtype=torch.float32).pin_memory()
        for o, t in slots:
            arena[o:o + t.numel()].copy_(t.reshape(-1))
        return arena, [(o, tuple(t.shape)) for o, t in slots], layout

    packed = [pack(v) for v in pin_views] if not args.no_e2e else None

    def stage_in(k):
        """H2D of everything one step reads, on the copy stream (overlaps the previous step's kernels)."""
        arena, slots, layout = packed[k]
        compute = torch.cuda.current_stream()
        with torch.cuda.stream(copy_stream):
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(copy_stream)
            d = arena.to(dev, non_blocking=True)
            d.record_stream(compute)
            c1.record(copy_stream)
            copy_events.append((c0, c1))
        get = lambda i: d[slots[i][0]:slots[i][0] + int(np.prod(slots[i][1]))].view(slots[i][1])
        view = [dict(G=st["G"], features=[get(i) for i in st["features"]], ref_proj=get(st["ref_proj"]),
                     src_projs=[get(i) for i in st["src_projs"]], logits=get(st["logits"]),
                     w=st["w"], eps=st["eps"]) for st in layout[:-1]]
        return (view, get(layout[-1])), c1

    H2s, W2s = syn.stage_shapes(h0, w0)[2]
    host_out = [[torch.empty((batch, h, w), dtype=torch.float32).pin_memory() for h, w in syn.stage_shapes(h0, w0)]
                + [torch.empty((batch, 2 * H2s, 2 * W2s), dtype=torch.float32).pin_memory()] for _ in range(2)]

    def e2e_loop(n):
        """Public-API calls with HOST (pinned) inputs: per step H2D of features / projections / depth_range / logits
        and D2H of the depth maps + confidence; the copies of step i+1 overlap the kernels of step i."""
        nxt = stage_in(0)
        for i in range(n):
            (view, drange), ev = nxt
            if i + 1 < n:
                nxt = stage_in((i + 1) % 2)
            torch.cuda.current_stream().wait_event(ev)
            depths, conf = hot_path_chain(view, drange)
            for dst, src in zip(host_out[i % 2], depths + [conf]):
                dst.copy_(src, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        # ---------------------------------------------------------------- device-resident throughput
        for i in range(max(args.warmup, 3)):
            hot_path(dev_views[i % 2])
        barrier()
        sampler = ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
        # (a) instrumented eager pass: CUDA events around every fused cost-volume call -> roofline numbers
        def make_events():
            evs = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
                   for _ in range(args.steps)]
            for step in evs:              # torch creates the CUDA event lazily: force it so the library gets a handle
                for a, b in step:
                    a.record(); b.record()
            return evs
        stage_events, kernel_events = make_events(), make_events()
        torch.cuda.synchronize()
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_start.record()
        for i in range(args.steps):
            hot_path(dev_views[i % 2], stage_events[i], kernel_events[i])
        t_end.record()
        barrier()
        eager_ms = t_start.elapsed_time(t_end)
        cv_ms = [statistics.mean(ev[s][0].elapsed_time(ev[s][1]) for ev in stage_events) for s in range(3)]
        hot_ms = [statistics.mean(ev[s][0].elapsed_time(ev[s][1]) for ev in kernel_events) for s in range(3)]

        # (b) the timed region: the same K steps, replayed from CUDA graphs (one per input set) so that the
        #     host's launch overhead is not on the critical path; falls back to eager launches if capture fails
        graphs, mode = None, "eager"
        if not args.no_graph:
            try:
                graphs = []
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    hot_path(dev_views[0]); hot_path(dev_views[1])
                torch.cuda.current_stream().wait_stream(side)
                keep = []
                for k in range(2):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        keep.append(hot_path(dev_views[k]))
                    graphs.append(g)
                for g in graphs:
                    g.replay()
                mode = "cuda-graph replay"
            except Exception as e:  # pragma: no cover
                graphs, mode = None, f"eager (graph capture failed: {type(e).__name__})"
                torch.cuda.synchronize()
        ops.reset_launch_count()
        barrier()
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_start.record()
        for i in range(args.steps):
            if graphs is not None:
                graphs[i % 2].replay()
            else:
                hot_path(dev_views[i % 2])
        t_end.record()
        barrier()
        launches = ops.launch_count() if graphs is None else args.steps * LAUNCHES_PER_STEP
        elapsed_ms = t_start.elapsed_time(t_end)

        # ------------------------------------------------------------------------- end to end (host)
        # The leg is PCIe bound (315 MB per step) and the host's PCIe / memory system is shared with the box's other
        # tenants: back-to-back runs were seen at 5.73, 5.73 and 8.68 ms per step.  So the K steps are timed E2E_PASSES
        # times, every pass is reported, and the fastest one is the figure (each pass: barrier, K steps, barrier;
        # max over ranks per pass).
        e2e_passes, h2d_gbps, bare_gbps = [], float("nan"), float("nan")
        if not args.no_e2e:
            e2e_loop(3)
            copy_events.clear()
            for _ in range(E2E_PASSES):
                e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                e_start.record()
                e2e_loop(args.steps)
                e_end.record()
                barrier()
                e2e_passes.append(e_start.elapsed_time(e_end))
            arena_bytes = packed[0][0].numel() * 4
            h2d_gbps = statistics.median(arena_bytes / 1e9 / (a.elapsed_time(b) / 1e3) for a, b in copy_events)
            # the limiter, measured: the same arena through a bare cudaMemcpyAsync, every rank at once, no kernels running
            barrier()
            bare = []
            for i in range(8):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); d = packed[i % 2][0].to(dev, non_blocking=True); b.record()
                torch.cuda.synchronize()
                bare.append(arena_bytes / 1e9 / (a.elapsed_time(b) / 1e3))
                del d
            bare_gbps = statistics.median(bare[2:])
            barrier()
        clocks = sampler.stop() if sampler else None

    per_rank = [[h2d_gbps, bare_gbps, float(bound_cpus or 0)]]
    if world > 1:
        t = torch.tensor([elapsed_ms] + (e2e_passes or [0.0]), device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t[0])
        e2e_passes = [float(x) for x in t[1:]] if e2e_passes else []
        mine = torch.tensor(per_rank[0], device=dev, dtype=torch.float64)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [[float(x) for x in r] for r in allr]
    e2e_ms = statistics.median(e2e_passes) if e2e_passes else float("nan")
    if world > 1:
        # the only exchange of the path: final depth + confidence maps to rank 0 (north_star), outside the hot loop
        depths, conf = hot_path(dev_views[0])
        gathered = [torch.empty_like(conf) for _ in range(world)] if rank == 0 else None
        dist.gather(conf, gathered, dst=0)
        torch.cuda.synchronize()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", FALLBACK_HBM_GBS))
        views = world * args.steps * batch
        h2d = 0
        if packed is not None:
            h2d = packed[0][0].numel() * 4          # what actually crosses PCIe per step: the arena incl. its alignment padding
        H2, W2 = syn.stage_shapes(h0, w0)[2]
        d2h = 4 * batch * (sum(h * w for h, w in syn.stage_shapes(h0, w0)) + 4 * H2 * W2)
        # the dominant kernel (cost_volume_staged_kernel: ~84 % of the step): algorithmic bytes of its three launches over
        # its own CUDA-event time; the op-level figure (layout pass + setup included) is kept next to it
        achieved = sum(cv_bytes) / 1e9 / (sum(hot_ms) / 1e3)
        achieved_op = sum(cv_bytes) / 1e9 / (sum(cv_ms) / 1e3)
        line = {
            "metric": "DTU views/s (plane-sweep cost-volume path)", "value": views / (elapsed_ms / 1e3), "unit": "views/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(workload),
            "e2e": None if args.no_e2e else {"value": views / (e2e_ms / 1e3), "unit": "views/s", "h2d_bytes_per_step": h2d,
                                             "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                                             "passes_ms_per_step": [t / args.steps for t in e2e_passes],
                                             "min_ms_per_step": min(e2e_passes) / args.steps, "best_pass_views_per_s": views / (min(e2e_passes) / 1e3),
                                             "h2d_GBps_per_rank": [r[0] for r in per_rank],
                                             "bare_memcpy_GBps_per_rank": [r[1] for r in per_rank],
                                             "cpus_bound_per_rank": [int(r[2]) for r in per_rank],
                                             "chain": "host sends features, projections, depth_range and logits; the hypotheses of stages 1-2 are formed on the "
                                                      "device (HyposByFit kernels fused into the head + mdf_hypos_generate_fwd, core.py:55 order)",
                                             "note": f"median of {E2E_PASSES} passes of K steps (max over ranks per pass), every pass listed; H2D-bound: "
                                                     "h2d_GBps_per_rank is the median rate of the per-step arena transfer inside the timed passes (CUDA events on "
                                                     "the copy stream), bare_memcpy_GBps_per_rank the same arena through cudaMemcpyAsync alone with every rank "
                                                     "copying at once"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (ncu_dram_traffic(workload) or {}).get("bytes"), "traffic_source": (ncu_dram_traffic(workload) or {}).get("source"),
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback",
                         "kernel": "cost_volume_staged_kernel<G=32|16|8> (the three launches of one step; events recorded inside the library around the kernel alone)",
                         "op_level": {"what": "mdf_cost_volume_fwd = setup_kernel || prep_kernel, then the kernel above", "GBps": achieved_op,
                                      "frac": achieved_op / peak},
                         "algorithmic_bytes_per_step": sum(cv_bytes),
                         "per_stage": [{"bytes": b, "ms": ms, "GB/s": b / 1e9 / (ms / 1e3), "hot_kernel_ms": hk}
                                       for b, ms, hk in zip(cv_bytes, cv_ms, hot_ms)],
                         "hot_kernel": "cost_volume_staged_kernel<G=32|16|8>: events recorded inside the library around the kernel alone",
                         "hot_kernel_ms_per_step": sum(hot_ms),
                         "hot_kernel_share_of_step": sum(hot_ms) / (eager_ms / args.steps),
                         "hot_kernel_GBps": sum(cv_bytes) / 1e9 / (sum(hot_ms) / 1e3),
                         "onchip_bound": onchip_bound(h0, w0, nviews, batch, (clocks or {}).get("sm_mhz") or 1965.0, sum(hot_ms)),
                         "timed_in": "instrumented eager pass of the same K steps (CUDA events around each call)",
                         "eager_ms_per_step": eager_ms / args.steps,
                         "share_of_step": sum(cv_ms) / (eager_ms / args.steps)},
            "launch_mode": mode,
            "clocks": clocks,
        }
        if world == 1 and not args.no_tail:
            try:
                line["regulariser_tail"] = time_regulariser_tail(dev, h0, w0, batch, host_views[0])
            except Exception as e:  # pragma: no cover - the headline must not depend on the extra
                line["regulariser_tail"] = {"error": f"{type(e).__name__}: {e}"}
        if world == 1 and not args.no_tail:
            try:
                del dev_views, pin_views, graphs
                torch.cuda.empty_cache()
                line["other_rows"] = time_other_rows(dev)
            except Exception as e:  # pragma: no cover - the headline must not depend on the extras
                line["other_rows"] = {"error": f"{type(e).__name__}: {e}"}
        if world == 1 and not args.no_tail:
            try:
                line["other_workloads"] = time_other_workloads(dev, [w for w in WORKLOADS if w != workload], peak)
            except Exception as e:  # pragma: no cover
                line["other_workloads"] = {"error": f"{type(e).__name__}: {e}"}
        if world == 1 and not args.no_reference:
            from oracle import ref_install
            if ref_install.available():
                from oracle import ref_bench
                try:
                    with torch.no_grad():
                        base = ref_bench.cuda_hot_path(host_views[0], dev)
                    base["speedup_device_resident"] = line["value"] / base["value"]
                    line["aten_cuda_baseline"] = base
                except Exception as e:  # pragma: no cover
                    line["aten_cuda_baseline"] = {"error": f"{type(e).__name__}: {e}"}
                try:
                    line["pipeline"] = ref_bench.pipeline(dev, h0, w0, nviews)
                except Exception as e:  # pragma: no cover
                    line["pipeline"] = {"error": f"{type(e).__name__}: {e}"}
            else:
                line["aten_cuda_baseline"] = line["pipeline"] = {"unavailable": "oracle/_ref did not travel with the snapshot (python -m oracle.ref_install)"}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(host_views[0], args.cpu_budget, batch)
        out.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


class quiet_stdout:
    """Everything libraries print to stdout while the benchmark runs (NCCL's version banner, ...) goes to stderr:
    stdout carries exactly one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line: str):
        sys.stdout.flush()
        os.write(self.saved, (line + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="dtu_1600x1152_n5")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of CUDA-graph replays")
    ap.add_argument("--no-tail", action="store_true", help="skip the extra timing of the fused regulariser tail, the other rows and workloads")
    ap.add_argument("--no-reference", action="store_true", help="skip the reference's modules on the same GPU (aten_cuda_baseline, pipeline)")
    args = ap.parse_args()
    with quiet_stdout() as out:
        if args.impl == "reference":
            run_reference(args, args.workload, out)
        else:
            run_b200(args, args.workload, out)


if __name__ == "__main__":
    main()
