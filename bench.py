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
}
FALLBACK_HBM_GBS = 6650.0      # /opt/skills/guides/B200_PROFILING.md fallback


# ----------------------------------------------------------------------------------------- workload
def make_view(h0, w0, nviews, batch, seed):
    """Host (numpy) inputs of one reference view: per stage features, projections, hypotheses,
    depth_weight parameters and regulariser logits."""
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
            logits=syn.regulariser_logits(batch, D, H, W, seed=seed + 40 + s)))
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
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in open(self.path).read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


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


def cpu_baseline(view, budget_s: float, batch: int):
    from oracle import c_oracle as co
    co.build()
    probe = cpu_sample(view, 1.0 / 16.0)                     # calibrate: 1/16 of the rows
    frac = min(1.0, max(1.0 / 16.0, budget_s / (probe * 16.0)))
    t = cpu_sample(view, frac)
    return {"value": batch * frac / t, "unit": "views/s", "cores": co.num_threads(), "kind": "port",
            "host_cpus": os.cpu_count(),
            "sample": f"first {frac:.3f} of the rows of all 3 stages of one view (cost volume + head), {t:.2f} s, "
                      f"C oracle oracle/mdf_oracle.c with OpenMP"}


def run_reference(args, workload):
    """--impl reference: the reference algorithm on the host cores (CPU oracle port; the reference itself is
    PyTorch-CPU Python that cannot travel to the GPU box).  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import c_oracle as co
    co.build()
    h0, w0, nviews, batch = WORKLOADS[workload]
    view = make_view(h0, w0, nviews, batch, seed=1)
    probe = cpu_sample(view, 1.0 / 16.0)
    total = max(1, args.steps + args.warmup)
    frac = min(1.0, max(1.0 / 32.0, (150.0 / total) / (probe * 16.0)))     # whole run within ~2.5 minutes
    for _ in range(args.warmup):
        cpu_sample(view, frac)
    times = [cpu_sample(view, frac) for _ in range(args.steps)]
    t = sum(times) / len(times)
    value = batch * frac / t
    sample = f"each step = first {frac:.3f} of the rows of all 3 stages of one view, {t:.2f} s/step"
    print(json.dumps({
        "impl": "reference", "metric": "DTU views/s (plane-sweep cost-volume path)", "value": value, "unit": "views/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / frac,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(workload),
        "cpu_baseline": {"value": value, "unit": "views/s", "cores": co.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ----------------------------------------------------------------------------------------- GPU arm
def run_b200(args, workload):
    import torch
    import torch.distributed as dist

    import mdf_net_b200 as mdf
    from mdf_net_b200 import ops

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this path has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
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
    pin_views = [convert(v, pin=True) for v in host_views] if not args.no_e2e else None

    def hot_path(view, events=None):
        """3 x (fused cost volume -> fused head).  Returns the per-stage depth maps and the confidence."""
        depths, conf = [], None
        for s, st in enumerate(view):
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

    def e2e_step(pview):
        """Public-API call with HOST (pinned) inputs: H2D of everything the path reads, D2H of its results."""
        view = []
        for st in pview:
            nb = lambda t: t.to(dev, non_blocking=True)
            view.append(dict(G=st["G"], features=[nb(f) for f in st["features"]], ref_proj=nb(st["ref_proj"]),
                             src_projs=[nb(q) for q in st["src_projs"]], hypos=nb(st["hypos"]), logits=nb(st["logits"]),
                             w=st["w"], eps=st["eps"]))
        depths, conf = hot_path(view)
        outs = [d.to("cpu", non_blocking=True) for d in depths] + [conf.to("cpu", non_blocking=True)]
        torch.cuda.current_stream().synchronize()
        return outs

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
        stage_events = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
                        for _ in range(args.steps)]
        ops.reset_launch_count()
        barrier()
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_start.record()
        for i in range(args.steps):
            hot_path(dev_views[i % 2], stage_events[i])
        t_end.record()
        barrier()
        launches = ops.launch_count()
        clocks = sampler.stop() if sampler else None
        elapsed_ms = t_start.elapsed_time(t_end)
        cv_ms = [statistics.mean(ev[s][0].elapsed_time(ev[s][1]) for ev in stage_events) for s in range(3)]

        # ------------------------------------------------------------------------- end to end (host)
        e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if not args.no_e2e:
            for i in range(3):
                e2e_step(pin_views[i % 2])
        barrier()
        e_start.record()
        if not args.no_e2e:
            for i in range(args.steps):
                e2e_step(pin_views[i % 2])
        e_end.record()
        barrier()
        e2e_ms = e_start.elapsed_time(e_end) if not args.no_e2e else float("nan")

    if world > 1:
        t = torch.tensor([elapsed_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_ms = float(t[0]), float(t[1])
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
        h2d = sum(sum(f.size for f in st["features"]) + st["ref_proj"].size + sum(q.size for q in st["src_projs"])
                  + st["hypos"].size + st["logits"].size for st in host_views[0]) * 4
        H2, W2 = syn.stage_shapes(h0, w0)[2]
        d2h = 4 * batch * (sum(h * w for h, w in syn.stage_shapes(h0, w0)) + 4 * H2 * W2)
        achieved = sum(cv_bytes) / 1e9 / (sum(cv_ms) / 1e3)
        line = {
            "metric": "DTU views/s (plane-sweep cost-volume path)", "value": views / (elapsed_ms / 1e3), "unit": "views/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(workload),
            "e2e": None if args.no_e2e else {"value": views / (e2e_ms / 1e3), "unit": "views/s", "h2d_bytes_per_step": h2d,
                                             "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback",
                         "kernel": "mdf_cost_volume_fwd (setup + prep + cost_volume_staged_kernel), 3 launches of the op per step",
                         "algorithmic_bytes_per_step": sum(cv_bytes),
                         "per_stage": [{"bytes": b, "ms": ms, "GB/s": b / 1e9 / (ms / 1e3)} for b, ms in zip(cv_bytes, cv_ms)],
                         "share_of_step": sum(cv_ms) / (elapsed_ms / args.steps)},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(host_views[0], args.cpu_budget, batch)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args, args.workload)
    else:
        run_b200(args, args.workload)


if __name__ == "__main__":
    main()
