#!/usr/bin/env python
"""Throughput bench of the VisTracker SIF-Net hot path on B200 (contract: task prompt + SURVEY.md section 8(d)).

Workload = BASELINE.json configs[1]: SIF-Net tri-vis-l2 forward, batch = 8 frames (512x512, 8 channels) + 10 000 query
points per frame, one GPU.  A "step" is ``filter(images)`` + one ``query(points)`` on the batch; the metric is frames/sec.
With N > 1 GPUs every rank runs the same batch shape on its own frames (frames of a sequence are independent: weak
scaling, no data-path collective).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

``--impl reference`` times the reference algorithm's CPU restatement (oracle/, PyTorch-CPU, all host threads) on a
bounded sample (1 frame + 10 000 points per step).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH, NPTS, SIZE = 8, 10000, 512
# conv-only 2*MAC per frame: 163.01 RGB encoder + 3 x 150.15 triplane encoder
GFLOP_FILTER_PER_FRAME = 613.46
# ncu dram__bytes_read.sum + dram__bytes_write.sum over the 186 tensor-core conv launches of one step, per launch (profiles/r01i_*)
CONV_DRAM_BYTES_PER_LAUNCH = 228.9e6
METRIC, UNIT = "frames/sec", "frames/s"
WORKLOAD = f"sifnet-tri-vis-l2 filter+query, batch={BATCH} frames 512x512x8ch, {NPTS} query points/frame"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_step(sd, frames, dims, n_points=NPTS, seed=100):
    """One pass of the CPU restatement (oracle/sifnet_ref.py) over `frames` frames; returns seconds."""
    import torch
    from oracle import sifnet_ref as R
    from vistracker_b200.synth import synthetic_frames
    images, points, crop, body = synthetic_frames(frames, size=SIZE, seed=seed, n_points=n_points, jitter=True)
    cam = (dims.fx_px, dims.fy_px, dims.cx_px, dims.cy_px, dims.crop_size)
    t0 = time.perf_counter()
    with torch.no_grad():
        maps = R.sif_filter(sd, images)
        R.sif_query(sd, maps, points, crop, body, cam)
    return time.perf_counter() - t0


def run_reference(args):
    """Reference arm: the reference algorithm on the host cores (oracle port -- the Python reference tree does not travel
    to the GPU box and its model classes are restated 1:1 in oracle/sifnet_ref.py, pinned by tests/golden)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from vistracker_b200 import default_options, resolve_dims
    from vistracker_b200.synth import synthetic_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dims = resolve_dims(default_options())
    sd = synthetic_state_dict(dims, seed=0)
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_step(sd, 1, dims)
    times = [cpu_oracle_step(sd, 1, dims) for _ in range(args.steps)]
    total = sum(times)
    v = args.steps / total
    sample = "1 frame 512x512x8ch + 10000 query points per step (filter+query), PyTorch-CPU fp32"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from vistracker_b200 import CHORETriplaneVisibility, default_options, resolve_dims
    from vistracker_b200 import encoder as enc_mod
    from vistracker_b200.synth import synthetic_frames, synthetic_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    dims = resolve_dims(default_options())
    sd = synthetic_state_dict(dims, seed=0)
    net = CHORETriplaneVisibility(default_options(), device=dev).eval()
    net.load_state_dict(sd)
    net.defer_checks = True

    images, points, crop, body = synthetic_frames(BATCH, size=SIZE, seed=2 + rank, n_points=NPTS, jitter=True)
    h_img, h_pts = images.pin_memory(), points.pin_memory()
    h_crop, h_body = crop.pin_memory(), body.pin_memory()
    d_img, d_pts, d_crop, d_body = (t.to(dev) for t in (images, points, crop, body))
    h_out = torch.empty(BATCH, 29, NPTS, dtype=torch.float32).pin_memory()

    def step_resident():
        net.filter(d_img)
        out, _ = net._query_raw(d_pts, d_crop, d_body)
        return out

    def step_e2e():
        net.filter(h_img.to(dev, non_blocking=True))
        net.query(h_pts.to(dev, non_blocking=True), crop_center=h_crop.to(dev, non_blocking=True),
                  body_center=h_body.to(dev, non_blocking=True))
        df, pca, parts, centers, vis = net.get_preds()
        h_out.copy_(df._base if df._base is not None else torch.cat([df, pca.flatten(1, 2), parts, centers, vis], 1), non_blocking=True)
        torch.cuda.current_stream().synchronize()       # the caller reads the predictions on the host every step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step_resident()
    net.check()
    sampler = ClockSampler(local) if rank == 0 else None
    ms = timed(step_resident, args.steps)
    clocks = sampler.stop() if sampler else None
    net.check()
    launches_per_step = net.launches_filter + 1
    value = world * BATCH * args.steps / (ms * 1e-3)

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e_value = world * BATCH * args.steps / (ms_e2e * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in (h_img, h_pts, h_crop, h_body))
    d2h = h_out.numel() * h_out.element_size()

    # ---- roofline of the dominant kernel (tcgen05 conv): events around every vt_conv_mma launch of one extra step
    roof = None
    if rank == 0:
        spans, conv_bytes = [], []
        orig_call = enc_mod._lib.call

        def traced(name, *a):
            if name not in ("vt_conv_mma", "vt_conv_mma_dual"):     # same leading arguments; _dual adds the fused residual output
                return orig_call(name, *a)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); orig_call(name, *a); e.record()
            n_img, H, W, cin_pad, _pad, ks, cout = a[2], a[3], a[4], a[5], a[6], a[7], a[10]
            # compulsory bytes of this launch: both fp16 operand planes and weight planes once, the fp32 output, residual reads and
            # (vt_conv_mma_dual) the second output + its residual
            n_out = 1 + (a[12] is not None) + (2 if name == "vt_conv_mma_dual" and a[18] is not None else 0)
            nbytes = (2 * n_img * (H + 2 * _pad) * (W + 2 * _pad) * cin_pad * 2 + 2 * ks * ks * cout * cin_pad * 2
                      + n_out * n_img * H * W * cout * 4)
            conv_bytes.append(nbytes)
            spans.append((s, e, 2.0 * n_img * H * W * cout * ks * ks, cin_pad, name))
            return None

        class _Shim:
            def __getattr__(self, k):
                return traced if k == "call" else getattr(orig_mod, k)
        orig_mod = enc_mod._lib
        enc_mod._lib = _Shim()
        try:
            # true input-channel counts (Cin, not the zero-padded Cin_pad) give the ALGORITHMIC flops
            cins = []
            orig_conv = enc_mod.HGEncoder._conv

            def conv_spy(self, op, name, out, **kw):
                before = len(spans)
                orig_conv(self, op, name, out, **kw)
                if len(spans) > before:
                    cins.append(op.act.C)
            enc_mod.HGEncoder._conv = conv_spy
            use_graph, net.use_graph = net.use_graph, False        # the traced pass launches eagerly so every conv can be bracketed
            net.filter(d_img)
            torch.cuda.synchronize()
        finally:
            enc_mod._lib = orig_mod
            enc_mod.HGEncoder._conv = orig_conv
            net.use_graph = use_graph
        t_ms = sum(s.elapsed_time(e) for s, e, *_ in spans)
        flops = sum(f * c for (_, _, f, _, _), c in zip(spans, cins))
        pk, src = peaks()
        achieved = flops / (t_ms * 1e-3) / 1e12 if spans else 0.0
        roof = {"bound": "tensor", "kernel": "conv_mma_persist_kernel (tcgen05 fp16x2-split: 3 MMA-equivalents per fp32 MAC)", "achieved": achieved,
                "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["bf16_tflops_sustained"],
                "traffic": CONV_DRAM_BYTES_PER_LAUNCH, "traffic_unit": "bytes per launch (ncu dram__bytes_read+write, average of the 186 launches of one step)",
                "traffic_source": "profiles/r01i_conv_dram_traffic.txt",
                "algorithmic_bytes_per_launch_avg": sum(conv_bytes) / max(len(conv_bytes), 1),
                "peak_source": f"{src} bf16 sustained (kernel timed inside a long step)",
                "executed_mma_frac": 3 * achieved / pk["bf16_tflops_sustained"], "launches": len(spans),
                "kernel_ms_per_step": t_ms, "share_of_step": t_ms / (ms / args.steps) if spans else 0.0,
                "algorithmic_gflop_per_launch_avg": flops / 1e9 / max(len(spans), 1)}

    if rank == 0:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cpu_frames = 2
        cpu_oracle_step(sd, 1, dims)                      # warm-up (thread pools, first-touch)
        t_cpu = cpu_oracle_step(sd, cpu_frames, dims)
        cpu = {"value": cpu_frames / t_cpu, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{cpu_frames} frames 512x512x8ch + {NPTS} points/frame, oracle/sifnet_ref.py (PyTorch-CPU fp32), {t_cpu:.1f} s"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_frames_per_step": world * BATCH, "parallelism": f"frame-parallel x{world}",
                       "l2": "inputs larger than L2: >2 GB of activations stream through HBM per step",
                       "conv_algo": os.environ.get("VT_CONV_ALGO", "mma"),
                       "filter_launch": "cuda-graph replay" if net.use_graph and net.filter_streams == 1 else "eager"},
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
            "filter_gflop_per_frame": GFLOP_FILTER_PER_FRAME,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
