#!/usr/bin/env python
"""Throughput bench of the VisTracker per-frame hot path on B200 (contract: task prompt + SURVEY.md section 8(d)).

Headline workload = BASELINE.json configs[3] ("C4"), the metric's own workload: the joint human-object optimisation of a sequence, measured on
its unit of work -- one 96-frame batch through ``recon_driver.fit_recon_batch``, the per-batch body of the reference's ``fit_recon``
(recon/recon_fit_triplane.py:47-111): neural reconstruction (SIF-Net filter + 40 projection steps on 20-30 k points per frame), filter of
the whole batch, ``optimize_smpl`` (reference schedule and early stop, <= 1030 Adam steps) and ``optimize_smpl_object`` (three phases,
<= 1550 steps).  A "step" is one such batch per GPU; a 1500-frame sequence is 16 of them.  frames/s = frames optimised / time.
With N > 1 GPUs every rank optimises its own batch (whole reference batches per rank, SURVEY.md 8(e): no data-path collective) and the
per-batch SMPL-T / object trajectories are stitched with ``parallel.gather_trajectory`` (NCCL all-gather) inside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c4|c2|c4seq] [--frames F]

``extra.c2`` keeps BASELINE config 2 (SIF-Net filter + 10 000-point query on 8 frames) with its own accuracy check; ``--workload c2`` makes
it the headline (round-1 line).  ``--workload c4seq`` times ONE pass over a fixed F-frame sequence sharded with ``parallel.rank_frames``
(strong scaling; run by hand, it takes minutes at N = 1).
``--impl reference`` times the reference algorithm's CPU restatement (oracle/, PyTorch-CPU, all host threads) on a bounded sample of the
same workload and extrapolates linearly (labelled so).
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

METRIC, UNIT = "frames/sec", "frames/s"
# ---- C4 (headline)
C4_FRAMES, SIZE = 96, 512
C4_WORKLOAD = ("recon_fit_trivis_full joint optimisation (C4): fit_recon_batch on 96 frames 512x512x8ch per GPU = neural reconstruction + filter + "
               "optimize_smpl (1+1+1+100 outer x 10) + optimize_smpl_object (15+30+<=110 outer x 10), reference schedules and early stops")
SMPL_CAPS = dict(iter_for_betas=1, iter_for_pose=1, iter_for_kpts=1, steps_per_iter=10, max_iter=100)          # recon_fit_triplane.py:66
OBJ_CAPS = dict(it_obj=15, it_sil=30, joint_iter=10, steps_per_iter=10, max_iter=100)                          # recon_fit_trivis_full.py:283-327
# bytes a point moves through the fused query-loss launch (forward gather 9 728 + second gather of the backward 9 728 +
# point 12 + label 8 + two values 8 + the merged point gradient 12)
QUERY_LOSS_BYTES_PER_POINT = 2 * 9728 + 12 + 8 + 8 + 12
NCU_QUERY_LOSS_DRAM_BYTES_PER_POINT = (543.143936e6 + 15.972608e6) / (96 * 6890)     # profiles/r02m_query_loss_merged_ncu_summary.txt
QUERY_LOSS_FLOP_PER_POINT = 2 * (1.117e6 / 5) * 3           # two heads, forward + the two backward products (1.117 MFLOP / 5 heads)
# ---- C2 (extra)
BATCH, NPTS = 8, 10000
GFLOP_FILTER_PER_FRAME = 613.46           # conv-only 2*MAC per frame: 163.01 RGB encoder + 3 x 150.15 triplane encoder
C2_WORKLOAD = f"sifnet-tri-vis-l2 filter+query, batch={BATCH} frames 512x512x8ch, {NPTS} query points/frame"


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
        # the loops alternate ~ms kernels with idle gaps: report the median of the samples taken under load (clock above idle)
        busy = [x for x in sm if x > 0.5 * max(mx)] if sm and mx else sm
        return {"sm_mhz": statistics.median(busy or sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# =====================================================================================================================================
# CPU side (oracle/): cpu_baseline leg and --impl reference only
# =====================================================================================================================================
# Weak scaling: every rank fits ITS OWN COPY of the same seeded batch.  The early stops of the optimisation are data dependent: with a different
# synthetic batch per rank (VT_BENCH_RANK_SEEDS=1: seed + 100 * rank) the max-over-ranks time measures which rank drew the batch with the longest
# joint phase (2.34 s at seed 4, 3.74 s at seed 104), not how the system scales.  extra.steps_taken_per_rank shows the balance either way.
RANK_SEEDS = os.environ.get("VT_BENCH_RANK_SEEDS", "0") == "1"
BODY = os.environ.get("VT_BENCH_BODY", "surface")             # "surface" (default) | "cloud" (the Gaussian point cloud of the earlier bench lines)
BODY_NOTE = {"surface": "human-shaped synthetic SMPL-H (synth_smpl.synthetic_smplh_surface: closed 1.7 m surface, vertices in surface order, proximity "
                        "skinning) -- SMPLH_male.pkl itself is licensed and absent",
             "cloud": "synth_smpl.synthetic_smplh: Gaussian point cloud in random vertex order (the body of the bench lines up to profiles/r02k)"}


def body_model(kind=None):
    """The SMPL-H stand-in of the C4 workload (same V = 6890, J = 52, buffers and work per vertex either way; only where the vertices lie differs)."""
    from vistracker_b200.synth_smpl import synthetic_smplh, synthetic_smplh_surface
    return synthetic_smplh_surface(seed=3) if (kind or BODY) == "surface" else synthetic_smplh(seed=3)


def c4_expected_steps():
    """Step counts of the seeded C4 batch as measured on the B200 (profiles/r02_c4_steps.json, written by a bench run) -- the CPU arm cannot
    know where the early stops fire without running hours of CPU optimisation; without the file the iteration caps are used."""
    caps = {"smpl": 1030, "object only": 150, "sil": 300, "joint": 1100, "source": "iteration caps (no early stop)"}
    p = os.path.join(ROOT, "profiles", "r02_c4_steps.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        d["source"] = "profiles/r02_c4_steps.json (steps the seeded batch took on the B200, early stops included)"
        return d
    return caps


def cpu_c4_sample(seed=4):
    """Per-unit CPU times of the C4 stages from the oracle restatements on a bounded sample, seconds:
    filter per frame; one generator projection step per frame (30 000 points, forward + gradient to the points); one optimize_smpl step per
    frame; one 'object only' step per frame; one 'joint' step per frame; one 'sil' step per frame (numpy rasteriser, 256 x 256, 1600 faces)."""
    import numpy as np
    import torch
    from oracle import recon_fit_ref as RF
    from oracle import sifnet_ref as SR
    from tools_inputs import load_assets
    from vistracker_b200 import default_options, resolve_dims
    from vistracker_b200.synth import synthetic_recon_batch, synthetic_state_dict
    n = 4                                                     # frames of the sample batch (the temporal terms need >= 4)
    a, reg = load_assets()
    d = synthetic_recon_batch(n, size=SIZE, seed=seed)
    sd = synthetic_state_dict(resolve_dims(default_options()), seed=0)
    model = body_model()
    t = {}
    t0 = time.perf_counter()
    with torch.no_grad():
        maps = SR.sif_filter(sd, d["images"][:2])
    t["filter_per_frame"] = (time.perf_counter() - t0) / 2
    with torch.no_grad():
        maps = SR.sif_filter(sd, d["images"])
    P = RF.Problem(sd, maps, model, reg, a, a["part_labels"].astype(np.int64), d["crop_center"], d["body_center"], net_in_size=SIZE)
    # generator: one projection step = query + backward to the points on 30 000 samples (recon/gen/generator.py:72-104), 1 frame
    P1 = RF.Problem(sd, {k: ([x[:1] for x in v] if isinstance(v, list) else v[:1]) for k, v in maps.items()}, model, reg, a,
                    a["part_labels"].astype(np.int64), d["crop_center"][:1], d["body_center"][:1], net_in_size=SIZE)
    pts = (d["body_center"][:1, None] + (torch.rand(1, 30000, 3) - 0.5) * torch.tensor([2.0, 3.0, 1.2])).requires_grad_(True)
    t0 = time.perf_counter()
    df = P1.query(pts)[0]
    torch.clamp(df[:, 0], max=2.0).sum().backward()
    t["generator_step_per_frame"] = time.perf_counter() - t0
    kp = d["body_kpts"]
    t0 = time.perf_counter()
    RF.optimize_smpl(P, d["pose"], d["betas"], d["trans"], d["pose"][:, 3:72].clone(), kp, 1, 1, 1, steps_per_iter=1, max_iter=0, step_budget=2)
    t["smpl_step_per_frame"] = (time.perf_counter() - t0) / (2 * n)
    rng = torch.Generator().manual_seed(seed)
    noise = lambda: torch.rand(n, 3, 3, generator=rng)
    keep, ref = torch.ones(n, 256, 256), d["images"][:, 4, ::2, ::2].contiguous()
    K = torch.tensor([[2.8, 0, 0.45], [0, 2.8, 0.5], [0, 0, 1]])[None].repeat(n, 1, 1)
    sil1 = RF.SilLoss(keep[:1], ref[:1], K[:1], d["obj_verts"].numpy(), d["obj_faces"].numpy(), rend_size=256)
    obj_t = d["body_center"] + torch.tensor([0.35, 0.0, 0.1])
    common = dict(objects=d["obj_points"][None].repeat(n, 1, 1), occ=d["occ_ratios"], noise_fn=noise, it_obj=1, it_sil=0, joint_iter=0, steps_per_iter=2, max_iter=1)
    t0 = time.perf_counter()
    out = RF.optimize_smpl_object(P, d["pose"], d["betas"], d["trans"], d["obj_rot_init"], obj_t, torch.ones(n), sil=None,
                                  step_budget={"object only": 2, "joint": 0}, **common)
    t["object_step_per_frame"] = (time.perf_counter() - t0) / (2 * n)
    t0 = time.perf_counter()
    RF.optimize_smpl_object(P, d["pose"], d["betas"], d["trans"], d["obj_rot_init"], obj_t, torch.ones(n), sil=None,
                            step_budget={"object only": 0, "joint": 2}, **common)
    t["joint_step_per_frame"] = (time.perf_counter() - t0) / (2 * n)
    R1 = RF.project_so3(d["obj_rot_init"][:1]).requires_grad_(True)
    t1 = obj_t[:1].clone().requires_grad_(True)
    t0 = time.perf_counter()
    sil1(R1, t1, torch.ones(1)).sum().backward()
    t["sil_step_per_frame"] = time.perf_counter() - t0
    return t


def cpu_c4_fps(unit, steps):
    """Linear extrapolation of the per-unit CPU times to one frame of the C4 batch: filter twice (generator mini-batches + whole batch,
    recon_fit_triplane.py:53-60), 2 targets x 2 rounds x 10 projection steps (20 000 points after the first round), the optimisation steps."""
    base = (2 * unit["filter_per_frame"] + 2 * (10 + 10 * 2 / 3) * unit["generator_step_per_frame"] + steps["smpl"] * unit["smpl_step_per_frame"]
            + steps["object only"] * unit["object_step_per_frame"] + steps["joint"] * unit["joint_step_per_frame"]
            + steps["sil"] * unit["object_step_per_frame"])              # a 'sil' step still transforms and queries the object points
    with_raster = base + steps["sil"] * unit["sil_step_per_frame"]
    # The reference has NO CPU implementation of its silhouette renderer (neural_renderer is CUDA-only); the numpy restatement under oracle/
    # is a checker (seconds per frame-step).  The baseline value therefore leaves the rasteriser's time OUT (conservative: a faster CPU arm);
    # the figure including it is reported next to it.
    return 1.0 / base, base, 1.0 / with_raster


def cpu_c2_step(sd, frames, dims, n_points=NPTS, seed=100, inputs=None):
    """One pass of the CPU restatement (oracle/sifnet_ref.py) over `frames` frames; returns (seconds, outputs)."""
    import torch
    from oracle import sifnet_ref as R
    from vistracker_b200.synth import synthetic_frames
    images, points, crop, body = inputs if inputs is not None else synthetic_frames(frames, size=SIZE, seed=seed, n_points=n_points, jitter=True)
    cam = (dims.fx_px, dims.fy_px, dims.cx_px, dims.cy_px, dims.crop_size)
    t0 = time.perf_counter()
    with torch.no_grad():
        maps = R.sif_filter(sd, images)
        out = R.sif_query(sd, maps, points, crop, body, cam)
    return time.perf_counter() - t0, out


def run_reference(args):
    """Reference arm: the reference algorithm on the host cores (oracle port -- the Python reference tree does not travel to the GPU box; its
    classes and loops are restated 1:1 under oracle/, pinned by tests/golden to the reference's own outputs)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from vistracker_b200 import default_options, resolve_dims
    from vistracker_b200.synth import synthetic_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if args.workload == "c2":
        dims = resolve_dims(default_options())
        sd = synthetic_state_dict(dims, seed=0)
        for _ in range(min(args.warmup, 1)):
            cpu_c2_step(sd, 1, dims)
        times = [cpu_c2_step(sd, 1, dims)[0] for _ in range(args.steps)]
        total = sum(times)
        v, ms_step = args.steps / total, 1e3 * total / args.steps
        workload = C2_WORKLOAD
        sample = "1 frame 512x512x8ch + 10000 query points per step (filter+query), PyTorch-CPU fp32"
        extra = {}
    else:
        steps = c4_expected_steps()
        cpu_c4_sample()                                       # warm-up (thread pools, first touch)
        units, t_all = [], 0.0
        for _ in range(max(1, min(args.steps, 4))):           # every sample is ~20 s of CPU work: at most 4 of them, the rest of --steps is not repeated
            t0 = time.perf_counter()
            units.append(cpu_c4_sample())
            t_all += time.perf_counter() - t0
        unit = {k: statistics.median(u[k] for u in units) for k in units[0]}
        v, per_frame, v_raster = cpu_c4_fps(unit, steps)
        ms_step = 1e3 * per_frame * C4_FRAMES
        workload = C4_WORKLOAD
        sample = (f"EXTRAPOLATED linearly from per-unit times of the oracle restatements on a 4-frame batch (filter 2 frames, 1 generator projection step "
                  f"on 30000 points, 2 optimize_smpl steps, 2 'object only' + 2 'joint' steps, 1 'sil' frame-step with the numpy rasteriser), "
                  f"{len(units)} samples of {t_all / len(units):.0f} s; step counts: {steps['source']}")
        extra = {"unit_seconds": unit, "steps_assumed": {k: steps[k] for k in ("smpl", "object only", "sil", "joint")}, "seconds_per_frame": per_frame,
                 "value_including_numpy_rasteriser": v_raster,
                 "note": "value excludes the silhouette rasteriser (no CPU implementation exists in the reference; the numpy checker would add sil_step_per_frame per 'sil' frame-step)"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": workload, "body_model": BODY_NOTE[BODY], "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, **extra},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# =====================================================================================================================================
# GPU side
# =====================================================================================================================================
class Dist:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = dist
        if self.world > 1:
            # NCCL writes its version banner (NCCL_DEBUG=VERSION in this image) to stdout: keep stdout for the ONE JSON line
            # (NCCL_DEBUG_FILE does not move the banner): drop the banner level, and point fd 1 at stderr while the communicator comes up
            if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=self.dev)
                dist.barrier()
                torch.cuda.synchronize()
            finally:
                os.dup2(saved, 1)
                os.close(saved)

    def barrier(self):
        import torch
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        """K calls bracketed by barrier + synchronize, CUDA events on the launching stream, MAX over ranks (ms); also the per-rank times."""
        import torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        per_rank = [ms]
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            allt = [torch.zeros_like(t) for _ in range(self.world)]
            self.dist.all_gather(allt, t)
            per_rank = [float(x.item()) for x in allt]
            ms = max(per_rank)
        return ms, per_rank

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


class C4:
    """The joint-optimisation batch of this rank: models, device / pinned-host inputs, the step functions."""

    def __init__(self, D: Dist, frames=C4_FRAMES, seed=4):
        import numpy as np
        import torch
        from tools_inputs import load_assets
        from vistracker_b200 import CHORETriplaneVisibility, default_options, resolve_dims
        from vistracker_b200.generator import GeneratorTriplaneVis
        from vistracker_b200.recon_fit import Priors, ReconFitterTriVisFull
        from vistracker_b200.smpl import LandmarkRegressor, SMPL_Layer
        from vistracker_b200.synth import synthetic_recon_batch, synthetic_state_dict
        dev = D.dev
        self.D, self.dev, self.frames = D, dev, frames
        a, reg = load_assets()
        self.dims = resolve_dims(default_options())
        self.sd = synthetic_state_dict(self.dims, seed=0)
        self.net = CHORETriplaneVisibility(default_options(), device=dev).eval()
        self.net.load_state_dict(self.sd)
        self.net.defer_checks = True
        model = body_model()
        self.layer = SMPL_Layer.from_buffers(model, model["parents"], dev)
        self.reg = LandmarkRegressor(np.stack([reg[0], reg[1]]), reg[2], reg[3], dev)
        h = synthetic_recon_batch(frames, size=SIZE, seed=seed + (100 * D.rank if RANK_SEEDS else 0))
        self.fitter = ReconFitterTriVisFull(self.net, Priors(a, dev), torch.from_numpy(a["part_labels"].astype(np.int64)),
                                            scan=(h["obj_verts"].numpy(), h["obj_faces"].numpy()))
        # random-init UDF: every in-front point counts as "on the surface" -> the minimum of 2 rounds per target a trained network needs
        self.gen = GeneratorTriplaneVis(self.net, threshold=2.0, filter_val=10.0)
        self.host = {k: v.pin_memory() for k, v in h.items()}
        self.devd = {k: v.to(dev) for k, v in h.items()}
        self.out_host = {k: torch.empty(*s).pin_memory() for k, s in (("pose", (frames, 156)), ("betas", (frames, 10)), ("trans", (frames, 3)),
                                                                      ("obj_R", (frames, 3, 3)), ("obj_t", (frames, 3)))}
        self.last = None

    def step(self, src, gather=True):
        """One batch through fit_recon_batch; ``src`` = self.devd (inputs resident in HBM) or self.host (pinned host buffers, copied inside)."""
        import torch
        from vistracker_b200 import parallel
        from vistracker_b200.pipeline import pack_neural
        from vistracker_b200.recon_driver import fit_recon_batch
        from vistracker_b200.recon_fit import SMPLParams
        dev = self.dev
        g = lambda k: src[k].to(dev, non_blocking=True)
        data = {"images": g("images"), "crop_center": g("crop_center"), "body_center": g("body_center")}
        pose, betas, trans = g("pose"), g("betas"), g("trans")
        init = lambda human_t: SMPLParams(self.layer, self.reg, pose, betas, trans)
        torch.manual_seed(1234)                                # the generator draws from torch's CPU generator, as the reference does
        out = fit_recon_batch(self.fitter, self.gen, data, init, g("body_kpts"), g("obj_points"), obj_rot_init=g("obj_rot_init"),
                              occ_ratios=g("occ_ratios"))
        smpl = out["smpl"]
        p = torch.cat([smpl.global_pose, smpl.body_pose, smpl.hand_pose], 1).detach()
        b = torch.cat([smpl.top_betas, smpl.other_betas], 1).detach()
        if gather and self.D.world > 1:
            # the one collective of the path: per-batch SMPL-T and object trajectories stitched for the sequence-global stages
            pc = out["pc_generated"]["object"]
            out["traj_smplt"] = parallel.gather_trajectory(parallel.pack_smplt(p, b, smpl.trans.detach()))
            out["traj_obj"] = parallel.gather_trajectory(pack_neural(out["obj_R"], out["obj_t"], pc["visibility"].to(dev)))
        if src is self.host:
            for k, v in (("pose", p), ("betas", b), ("trans", smpl.trans.detach()), ("obj_R", out["obj_R"]), ("obj_t", out["obj_t"])):
                self.out_host[k].copy_(v, non_blocking=True)
            torch.cuda.current_stream().synchronize()          # the caller reads the parameters on the host (save_outputs)
        self.last = out
        return out

    def counts(self, out):
        """Optimisation steps of the last batch per phase and kernel launches."""
        f = self.fitter
        it_obj, it_sil = f.get_opt_iters()["object"], f.get_opt_iters()["sil"]
        n_obj = len(out["hist_obj"])
        return {"smpl": len(out["hist_smpl"]), "object only": min(n_obj, it_obj * 10), "sil": max(0, min(n_obj - it_obj * 10, it_sil * 10)),
                "joint": max(0, n_obj - (it_obj + it_sil) * 10), "stopped_smpl": bool(out["stopped_smpl"]), "stopped_obj": bool(out["stopped_obj"])}

    def h2d_bytes(self):
        return sum(v.numel() * v.element_size() for k, v in self.host.items() if k not in ("obj_verts", "obj_faces"))

    def d2h_bytes(self):
        return sum(v.numel() * v.element_size() for v in self.out_host.values())


def query_roofline(c4: C4, share_launches, step_ms):
    """The dominant kernel of the metric's workload: query_bwd_tc_kernel in its fused-loss mode (vt_query_losses_tc) on the 96 x 6890 SMPL
    vertices -- one launch per optimize_smpl step.  Timed with CUDA events over repeated launches on the batch's own maps and vertices
    (inside the CUDA-graph steps a single kernel cannot be bracketed)."""
    import torch
    net, dev = c4.net, c4.dev
    B, V = c4.frames, c4.layer.V
    with torch.no_grad():
        verts = c4.last["smpl"]()[0].detach().contiguous()
    f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    labels = c4.fitter.part_labels.to(dev)[None].repeat(B, 1).contiguous()
    vals_df, g_df, vals_ce = f(B, V), f(B, V, 3), f(B, V)
    cc, bc = c4.devd["crop_center"].contiguous(), c4.devd["body_center"].contiguous()
    w = torch.tensor([30.0 ** 2, 0.05 ** 2], device=dev)   # the launch the optimize_smpl step replays: both heads merged, weights from device words
    run = lambda: net.enqueue_query_losses_merged(verts, cc, bc, 0, 0.1, labels, w.data_ptr(), 1.0 / (B * V), w.data_ptr() + 4, 1.0 / B,
                                                  vals_df, vals_ce, g_df)
    n = 20

    def timed_launches():
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            run()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    ms = timed_launches()
    pk, src = peaks()
    nbytes = QUERY_LOSS_BYTES_PER_POINT * B * V
    achieved = nbytes / (ms * 1e-3) / 1e9
    # the same launch on the OTHER synthetic body at the batch's initial parameters (continuity with the earlier bench lines: where the vertices
    # lie decides how well the taps of a 128-point tile share cache lines, nothing else differs)
    other = "cloud" if BODY == "surface" else "surface"
    try:
        from vistracker_b200.smpl import SMPL_Layer
        mo = body_model(other)
        lo = SMPL_Layer.from_buffers(mo, mo["parents"], dev)
        with torch.no_grad():
            vo = lo(c4.devd["pose"], c4.devd["betas"], c4.devd["trans"])[0].detach().contiguous()
        verts.copy_(vo)
        ms_o = timed_launches()
        other_body = {"body": BODY_NOTE[other], "ms_per_launch": ms_o, "frac": nbytes / (ms_o * 1e-3) / 1e9 / pk["hbm_gbs"]}
    except Exception as ex:              # noqa: BLE001
        other_body = {"error": f"{type(ex).__name__}: {ex}"}
    return {"bound": "hbm", "kernel": "query_bwd_tc_kernel, fused-loss mode with merged heads (vt_query_losses_merged_tc) on 96 x 6890 vertices: 1 launch per optimize_smpl step",
            "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
            "traffic": NCU_QUERY_LOSS_DRAM_BYTES_PER_POINT * B * V,
            "traffic_note": "dram__bytes_read + write of one ncu --set full capture of this launch (profiles/r02m_query_loss_merged_ncu_summary.txt: 559.1 MB at 96 x 6890 points of a surface body), scaled per point; "
                            "far BELOW the algorithmic bytes: the bilinear taps are L1- (67 %) and L2-served (91 %), the 8 maps of a frame are 71 MB and the vertices of one body touch a small part",
            "algorithmic_bytes_per_launch": nbytes, "bytes_per_point": QUERY_LOSS_BYTES_PER_POINT, "ms_per_launch": ms, "body": BODY_NOTE[BODY], "other_body": other_body,
            "tensor_tflops": QUERY_LOSS_FLOP_PER_POINT * B * V / (ms * 1e-3) / 1e12, "launches_per_step": share_launches,
            "share_of_step": share_launches * ms / step_ms, "peak_source": f"{src} HBM copy bandwidth",
            "timing": "CUDA events around 20 back-to-back launches on the launching stream after the timed region"}


def c2_extra(D: Dist, net, sd, dims, steps):
    """BASELINE config 2 (round-1 headline) kept as an extra: device-resident and end-to-end frames/s, the tensor-core conv roofline, and the
    accuracy of frame 0 of ITS OWN batch against the CPU oracle."""
    import torch
    from vistracker_b200 import encoder as enc_mod
    from vistracker_b200.synth import synthetic_frames
    dev = D.dev
    images, points, crop, body = synthetic_frames(BATCH, size=SIZE, seed=2 + D.rank, n_points=NPTS, jitter=True)
    h = [t.pin_memory() for t in (images, points, crop, body)]
    d_img, d_pts, d_crop, d_body = (t.to(dev) for t in (images, points, crop, body))
    h_out = torch.empty(BATCH, 29, NPTS, dtype=torch.float32).pin_memory()

    def step_resident():
        net.filter(d_img)
        return net._query_raw(d_pts, d_crop, d_body)[0]

    def step_e2e():
        net.filter(h[0].to(dev, non_blocking=True))
        out, _ = net._query_raw(h[1].to(dev, non_blocking=True), h[2].to(dev, non_blocking=True), h[3].to(dev, non_blocking=True))
        h_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    for _ in range(3):
        step_resident()
    ms, _ = D.timed(step_resident, steps)
    step_e2e()
    ms_e2e, _ = D.timed(step_e2e, steps)
    res = {"workload": C2_WORKLOAD, "value": D.world * BATCH * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps,
           "e2e": {"value": D.world * BATCH * steps / (ms_e2e * 1e-3), "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in h),
                   "d2h_bytes_per_step": h_out.numel() * 4}, "steps": steps, "gpu_launches_per_step": net.launches_filter + 1}
    if D.rank != 0:
        return res
    # conv roofline: events around every tcgen05 conv launch of one eager pass
    spans, cins = [], []
    orig_mod, orig_conv = enc_mod._lib, enc_mod.HGEncoder._conv

    def traced(name, *a):
        if name not in ("vt_conv_mma", "vt_conv_mma_dual"):
            return orig_mod.call(name, *a)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); orig_mod.call(name, *a); e.record()
        spans.append((s, e, 2.0 * a[2] * a[3] * a[4] * a[10] * a[7] * a[7]))

    class _Shim:
        def __getattr__(self, k):
            return traced if k == "call" else getattr(orig_mod, k)

    def conv_spy(self, op, name, out, **kw):
        before = len(spans)
        orig_conv(self, op, name, out, **kw)
        if len(spans) > before:
            cins.append(op.act.C)
    enc_mod._lib, enc_mod.HGEncoder._conv = _Shim(), conv_spy
    use_graph, net.use_graph = net.use_graph, False
    try:
        net.filter(d_img)
        torch.cuda.synchronize()
    finally:
        enc_mod._lib, enc_mod.HGEncoder._conv, net.use_graph = orig_mod, orig_conv, use_graph
    t_ms = sum(s.elapsed_time(e) for s, e, _ in spans)
    flops = sum(f * c for (_, _, f), c in zip(spans, cins))
    pk, src = peaks()
    ach = flops / (t_ms * 1e-3) / 1e12 if spans else 0.0
    res["roofline"] = {"bound": "tensor", "kernel": "conv_mma_persist_kernel (tcgen05, fp16 hi/lo split: 3 MMA-equivalents per fp32 MAC -> algorithmic ceiling 0.333)",
                       "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops_sustained"],
                       "executed_mma_frac": 3 * ach / pk["bf16_tflops_sustained"], "launches": len(spans), "kernel_ms_per_step": t_ms,
                       "share_of_step": t_ms / (ms / steps), "peak_source": f"{src} bf16 sustained"}
    # accuracy: frame 0 of this very batch through the CPU oracle
    net.filter(d_img)
    ours = net._query_raw(d_pts, d_crop, d_body)[0][0].cpu()
    torch.set_num_threads(os.cpu_count() or 1)
    _, ref = cpu_c2_step(sd, 1, dims, inputs=(images[:1], points[:1], crop[:1], body[:1]))
    ref = torch.cat([ref[0][0], ref[1][0].reshape(9, -1), ref[2][0], ref[3][0], ref[4][0]], 0)
    names, sl = ("df", "pca", "parts", "centers", "visibility"), ((0, 2), (2, 11), (11, 25), (25, 28), (28, 29))
    acc = {}
    for nme, (lo, hi) in zip(names, sl):
        a_, b_ = ours[lo:hi].double(), ref[lo:hi].double()
        acc[nme] = {"max_rel": float((a_ - b_).abs().max() / b_.abs().max()),
                    "max_elementwise_rel": float(((a_ - b_).abs() / torch.clamp(b_.abs(), min=1e-3 * float(b_.abs().max()))).max())}
    res["accuracy"] = {"vs": "oracle/sifnet_ref.py (PyTorch-CPU fp32) on frame 0 of the benched batch", "heads": acc,
                       "max_rel": max(v["max_rel"] for v in acc.values()), "max_elementwise_rel": max(v["max_elementwise_rel"] for v in acc.values()),
                       "elementwise_rule": "|a-b| <= tol * max(|b|, 1e-3 * max|b|) (SURVEY.md section 7)"}
    return res


def torch_eager_gpu(D: Dist, sd, dims, steps=3):
    """R-GPU context number: the reference's network restated in plain PyTorch (oracle/sifnet_ref.py: F.conv2d /
    group_norm / grid_sample -> cuDNN + ATen kernels) on the same B200, fp32 with TF32 off (the reference's torch 1.6 behaviour) and on."""
    import torch
    from oracle import sifnet_ref as R
    from vistracker_b200.synth import synthetic_frames
    dev = D.dev
    sdg = {k: v.to(dev) for k, v in sd.items()}
    images, points, crop, body = (t.to(dev) for t in synthetic_frames(BATCH, size=SIZE, seed=2, n_points=NPTS, jitter=True))
    cam = (dims.fx_px, dims.fy_px, dims.cx_px, dims.cy_px, dims.crop_size)
    out = {}
    keep = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = tf32

        def step():
            with torch.no_grad():
                maps = R.sif_filter(sdg, images)
                R.sif_query(sdg, maps, points, crop, body, cam)
        step(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record(); torch.cuda.synchronize()
        out["tf32_on" if tf32 else "tf32_off"] = {"frames_per_s": BATCH * steps / (e0.elapsed_time(e1) * 1e-3), "ms_per_step": e0.elapsed_time(e1) / steps}
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = keep
    out["what"] = "oracle/sifnet_ref.py on cuda (cuDNN / ATen eager), C2 shape; a baseline, none of this repo's kernels"
    return out


def torch_eager_gpu_c4(dev, counts, frames=C4_FRAMES, size=SIZE, gen_points=30000, gen_frames=16, n_steps=3, seed=4):
    """R-GPU context number for the metric's own workload: the reference's algorithm of every C4 stage
    as PLAIN PyTorch on the same B200 -- the restatements under oracle/ (pinned to the reference's own loops) run under ``torch.device(cuda)``:
    cuDNN / ATen eager kernels, autograd, torch.optim.Adam, fp32 with TF32 off (torch 1.6 behaviour), none of this repo's kernels.  Per-unit times
    (filter per frame on a 16-frame mini-batch; one generator projection step on 16 x 30 000 points; optimize_smpl / 'object only' / 'joint'
    steps on the 96-frame batch), extrapolated to the step counts of the batch exactly like the CPU baseline (cpu_c4_fps); the silhouette
    rasteriser is left out of the value on both sides of that formula (the restatement has no GPU rasteriser, the reference's is neural_renderer)."""
    import numpy as np
    import torch
    from oracle import recon_fit_ref as RF
    from oracle import sifnet_ref as SR
    from tools_inputs import load_assets
    from vistracker_b200 import default_options, resolve_dims
    from vistracker_b200.synth import synthetic_recon_batch, synthetic_state_dict
    dev = torch.device(dev)
    on_gpu = dev.type == "cuda"
    sync = (lambda: torch.cuda.synchronize(dev)) if on_gpu else (lambda: None)
    keep = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    a, reg = load_assets()
    d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in synthetic_recon_batch(frames, size=size, seed=seed).items()}
    sd = {k: v.to(dev) for k, v in synthetic_state_dict(resolve_dims(default_options()), seed=0).items()}
    model = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in body_model().items()}
    labels = a["part_labels"].astype(np.int64)
    t = {}

    def timed(fn, n=1):
        fn(); sync()                                            # warm-up (cuDNN algorithm selection, allocator)
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        sync()
        return (time.perf_counter() - t0) / n
    try:
        with torch.device(dev):
            gf = min(gen_frames, frames)
            with torch.no_grad():
                t["filter_per_frame"] = timed(lambda: SR.sif_filter(sd, d["images"][:gf])) / gf
                chunks = [SR.sif_filter(sd, d["images"][i:i + gf]) for i in range(0, frames, gf)]
            maps = {k: ([torch.cat([c[k][j] for c in chunks]) for j in range(len(chunks[0][k]))] if isinstance(chunks[0][k], (list, tuple))
                        else torch.cat([c[k] for c in chunks])) for k in chunks[0]}
            del chunks
            P = RF.Problem(sd, maps, model, reg, a, labels, d["crop_center"], d["body_center"], net_in_size=size)
            # generator: one projection step = query + backward to the points (recon/gen/generator.py:72-104) on a mini-batch
            P1 = RF.Problem(sd, {k: ([x[:gf] for x in v] if isinstance(v, (list, tuple)) else v[:gf]) for k, v in maps.items()}, model, reg, a, labels,
                            d["crop_center"][:gf], d["body_center"][:gf], net_in_size=size)
            g = torch.Generator(device="cpu").manual_seed(seed)
            pts0 = d["body_center"][:gf, None] + ((torch.rand(gf, gen_points, 3, generator=g, device="cpu") - 0.5) * torch.tensor([2.0, 3.0, 1.2], device="cpu")).to(dev)

            def gen_step():
                pts = pts0.clone().requires_grad_(True)
                torch.clamp(P1.query(pts)[0][:, 0], max=2.0).sum().backward()
            t["generator_step_per_frame"] = timed(gen_step, 2) / gf
            kp, pose_init = d["body_kpts"], d["pose"][:, 3:72].clone()
            t["smpl_step_per_frame"] = timed(lambda: RF.optimize_smpl(P, d["pose"], d["betas"], d["trans"], pose_init, kp, 1, 1, 1, steps_per_iter=1, max_iter=0,
                                                                     step_budget=n_steps)) / (n_steps * frames)
            rng = torch.Generator(device="cpu").manual_seed(seed)
            noise = lambda: torch.rand(frames, 3, 3, generator=rng, device="cpu").to(dev)
            obj_t = d["body_center"] + torch.tensor([0.35, 0.0, 0.1])
            common = dict(objects=d["obj_points"][None].repeat(frames, 1, 1), occ=d["occ_ratios"], noise_fn=noise, it_obj=1, it_sil=0, joint_iter=0,
                          steps_per_iter=n_steps, max_iter=1)
            for key, budget in (("object_step_per_frame", {"object only": n_steps, "joint": 0}), ("joint_step_per_frame", {"object only": 0, "joint": n_steps})):
                t[key] = timed(lambda: RF.optimize_smpl_object(P, d["pose"], d["betas"], d["trans"], d["obj_rot_init"], obj_t, torch.ones(frames), sil=None,
                                                               step_budget=budget, **common)) / (n_steps * frames)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = keep
    t["sil_step_per_frame"] = float("nan")
    fps, per_frame, _ = cpu_c4_fps(t, counts)
    return {"frames_per_s": fps, "seconds_per_frame": per_frame, "unit_seconds_per_frame": {k: v for k, v in t.items() if v == v},
            "ms_per_step_on_the_batch": {"optimize_smpl": t["smpl_step_per_frame"] * frames * 1e3, "object only": t["object_step_per_frame"] * frames * 1e3,
                                         "joint": t["joint_step_per_frame"] * frames * 1e3, "generator projection (16 frames x 30000 points)": t["generator_step_per_frame"] * gf * 1e3},
            "what": "oracle/ restatements of the reference's C4 stages as plain PyTorch on cuda (cuDNN / ATen eager + autograd + torch.optim.Adam, fp32, TF32 off), per-unit times "
                    "EXTRAPOLATED to this batch's step counts with the CPU baseline's formula (silhouette rasteriser excluded); a baseline, none of this repo's kernels"}


def c4_accuracy(D: Dist):
    """The accuracy half of the metric on a problem the CPU can finish: the same 4-frame batch (64 x 64 network input) through both loops on
    the GPU and through the CPU oracle (oracle/recon_fit_ref.py, pinned to the reference's own loops) with the same decopose_axis draws;
    parameters compared with the 1e-4 tolerance form, meshes with the evaluation's Chamfer distance (recon/eval/chamfer_distance.py via
    oracle/geom_ref.eval_chamfer on the vertices / object points, in cm)."""
    import numpy as np
    import torch
    from oracle import geom_ref as GR
    from oracle import recon_fit_ref as RF
    from oracle import sifnet_ref as SR
    from tools_inputs import load_assets
    from vistracker_b200 import CHORETriplaneVisibility, default_options, resolve_dims
    from vistracker_b200.recon_fit import Priors, ReconFitterTriVisFull, SMPLParams
    from vistracker_b200.render import SilLossROI
    from vistracker_b200.smpl import LandmarkRegressor, SMPL_Layer
    from vistracker_b200.synth import synthetic_recon_batch, synthetic_state_dict
    dev, n, S = D.dev, 4, 64
    a, reg = load_assets()
    dims = resolve_dims(default_options())
    sd = synthetic_state_dict(dims, seed=0)
    model = body_model()
    d = synthetic_recon_batch(n, size=S, seed=9, n_obj_points=600, obj_rings=6, obj_segments=8)
    net = CHORETriplaneVisibility(default_options(), device=dev).eval()
    net.load_state_dict(sd)
    net.filter(d["images"].to(dev))
    layer = SMPL_Layer.from_buffers(model, model["parents"], dev)
    body25 = LandmarkRegressor(np.stack([reg[0], reg[1]]), reg[2], reg[3], dev)
    fitter = ReconFitterTriVisFull(net, Priors(a, dev), torch.from_numpy(a["part_labels"].astype(np.int64)), net_in_size=S)
    c = lambda t: t.to(dev)
    qd = {"crop_center": c(d["crop_center"]), "body_center": c(d["body_center"])}
    kw_s = dict(iter_for_betas=1, iter_for_pose=1, iter_for_kpts=1, steps_per_iter=2, max_iter=4)
    smpl = SMPLParams(layer, body25, d["pose"], d["betas"], d["trans"])
    pose_init = d["pose"][:, 3:72].clone()
    dd = {"part_labels": c(torch.from_numpy(a["part_labels"].astype(np.int64)))[None].repeat(n, 1), "query_dict": qd, "pose_init": c(pose_init),
          "body_kpts": c(d["body_kpts"])}
    smpl, _ = fitter.optimize_smpl(smpl, dd, **kw_s)
    hist_s = np.asarray(fitter.last_hist)
    keep, ref = torch.ones(n, S, S), d["images"][:, 4].contiguous()
    K = torch.tensor([[2.8, 0, 0.45], [0, 2.8, 0.5], [0, 0, 1]])[None].repeat(n, 1, 1)
    sil = SilLossROI(keep, ref, K, d["obj_verts"].numpy(), d["obj_faces"].numpy(), rend_size=S, device=dev)
    noise_seq = torch.rand(64, n, 3, 3, generator=torch.Generator().manual_seed(5))
    obj_t0 = d["body_center"] + torch.tensor([0.35, 0.0, 0.1])
    draws = [0]

    def noise_gpu():
        draws[0] += 1
        return noise_seq[draws[0] - 1].to(dev)
    od = {"smpl": smpl, "query_dict": qd, "obj_R": c(d["obj_rot_init"]).clone().requires_grad_(True), "obj_t": c(obj_t0).clone().requires_grad_(True),
          "obj_s": torch.ones(n, device=dev), "objects": c(d["obj_points"])[None].repeat(n, 1, 1).contiguous(), "occ_ratios": c(d["occ_ratios"]), "silhouette": sil}
    fitter.get_opt_iters = staticmethod(lambda: {"sil": 1, "object": 2})
    _, R_g, t_g = fitter.optimize_smpl_object(net, od, joint_iter=1, steps_per_iter=2, max_iter=4, noise_fn=noise_gpu)
    hist_o = np.asarray(fitter.last_hist)
    R_gf = fitter.final_rotation(R_g)
    with torch.no_grad():
        verts_g = smpl()[0].cpu()
        obj_g = fitter.transform_obj_verts(od["objects"], R_gf, t_g.detach(), od["obj_s"]).cpu()
    # ---- the same on the CPU oracle
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        maps = SR.sif_filter(sd, d["images"])
    P = RF.Problem(sd, maps, model, reg, a, a["part_labels"].astype(np.int64), d["crop_center"], d["body_center"], net_in_size=S)
    rs = RF.optimize_smpl(P, d["pose"], d["betas"], d["trans"], pose_init, d["body_kpts"], **kw_s)
    sil_c = RF.SilLoss(keep, ref, K, d["obj_verts"].numpy(), d["obj_faces"].numpy(), rend_size=S)
    cd = [0]

    def noise_cpu():
        cd[0] += 1
        return noise_seq[cd[0] - 1]
    ro = RF.optimize_smpl_object(P, rs["pose"], rs["betas"], rs["trans"], d["obj_rot_init"], obj_t0, torch.ones(n), d["obj_points"][None].repeat(n, 1, 1),
                                 d["occ_ratios"], sil_c, noise_cpu, it_obj=2, it_sil=1, joint_iter=1, steps_per_iter=2, max_iter=4)
    with torch.no_grad():
        verts_c = P.smpl(rs["pose"], rs["betas"], rs["trans"])
        obj_c = GR.transform_obj_verts(d["obj_points"][None].repeat(n, 1, 1), ro["rot_final"], ro["obj_t"], torch.ones(n))
    rel = lambda x, y: float((x.double() - y.double()).abs().max() / y.double().abs().max())
    pose_g = torch.cat([smpl.global_pose, smpl.body_pose, smpl.hand_pose], 1).detach().cpu()
    betas_g = torch.cat([smpl.top_betas, smpl.other_betas], 1).detach().cpu()
    cham_s = [100 * GR.eval_chamfer(verts_g[i, ::4].numpy(), verts_c[i, ::4].numpy()) for i in range(n)]
    cham_o = [100 * GR.eval_chamfer(obj_g[i].numpy(), obj_c[i].numpy()) for i in range(n)]
    same_len = len(hist_s) == len(rs["hist"]) and len(hist_o) == len(ro["hist"])
    return {"vs": "oracle/recon_fit_ref.py (CPU restatement pinned to the reference's own optimize_smpl / optimize_smpl_object, tests/golden/recon_*loop.npz)",
            "problem": f"{n} frames, {S}x{S} network input, {len(rs['hist'])} optimize_smpl steps + {len(ro['hist'])} object steps (all three phases), same noise draws",
            "same_step_counts": same_len,
            "max_rel": {"pose": rel(pose_g, rs["pose"]), "betas": rel(betas_g, rs["betas"]), "trans": rel(smpl.trans.detach().cpu(), rs["trans"]),
                        "obj_rot": rel(R_gf.cpu(), ro["rot_final"]), "obj_trans": rel(t_g.detach().cpu(), ro["obj_t"]),
                        "loss_history_smpl": rel(torch.from_numpy(hist_s[:len(rs["hist"])]), torch.from_numpy(rs["hist"][:len(hist_s)])),
                        "loss_history_obj": rel(torch.from_numpy(hist_o[:len(ro["hist"])]), torch.from_numpy(ro["hist"][:len(hist_o)]))},
            "chamfer_cm": {"smpl": float(np.mean(cham_s)), "object": float(np.mean(cham_o)),
                           "how": "recon/eval/chamfer_distance.py (bidirectional mean NN distance x 100) between this implementation's and the oracle's result meshes"}}


def run_ours(args):
    import torch
    D = Dist()
    warmup = max(args.warmup, 3)
    if args.workload == "c2":
        from vistracker_b200 import CHORETriplaneVisibility, default_options, resolve_dims
        from vistracker_b200.synth import synthetic_state_dict
        dims = resolve_dims(default_options())
        sd = synthetic_state_dict(dims, seed=0)
        net = CHORETriplaneVisibility(default_options(), device=D.dev).eval()
        net.load_state_dict(sd)
        net.defer_checks = True
        sampler = ClockSampler(D.local) if D.rank == 0 else None
        res = c2_extra(D, net, sd, dims, args.steps)
        clocks = sampler.stop() if sampler else None
        if D.rank == 0:
            print(json.dumps({"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": D.world, "steps": args.steps, "warmup": 3,
                              "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                              "data": "synthetic", "config": {"workload": C2_WORKLOAD, "l2": "inputs larger than L2"}, "roofline": res.get("roofline"),
                              "accuracy": res.get("accuracy"), "e2e": {**res["e2e"], "unit": UNIT}, "gpu_launches": res["gpu_launches_per_step"] * args.steps,
                              "clocks": clocks}))
        D.close()
        return
    if args.workload == "c4seq":
        return run_sequence(D, args)

    c4 = C4(D)
    for _ in range(warmup):
        c4.step(c4.devd)
    c4.net.check()
    sampler = ClockSampler(D.local) if D.rank == 0 else None
    ms, per_rank = D.timed(lambda: c4.step(c4.devd), args.steps)
    clocks = sampler.stop() if sampler else None
    c4.net.check()
    counts = c4.counts(c4.last)
    value = D.world * c4.frames * args.steps / (ms * 1e-3)
    e2e_steps = max(1, min(args.steps, 5))
    c4.step(c4.host)
    ms_e2e, _ = D.timed(lambda: c4.step(c4.host), e2e_steps)
    e2e_value = D.world * c4.frames * e2e_steps / (ms_e2e * 1e-3)
    step_ms = ms / args.steps
    # launches of this repo's kernels per batch: 16 per optimize_smpl step, ~12-19 per object step, ~440 per filter call (12 calls: 6 generator
    # mini-batches + 6 chunks of the whole-batch filter), 1 per generator projection step (2 targets x 2 rounds x 10 x 6 mini-batches) + forward queries
    counts_all = [counts]
    if D.world > 1:
        counts_all = [None] * D.world
        D.dist.all_gather_object(counts_all, counts)
    launches = (counts["smpl"] * 16 + counts["object only"] * 11 + counts["sil"] * 16 + counts["joint"] * 17 + 12 * c4.net.launches_filter + 6 * 2 * 2 * 11)
    if D.rank == 0:
        roof = query_roofline(c4, counts["smpl"], step_ms)
        extra = {"per_rank_ms": [x / args.steps for x in per_rank], "steps_taken": counts, "steps_taken_per_rank": counts_all,
                 "rank_batches": "a different seeded batch per rank" if RANK_SEEDS else "every rank fits its own copy of the same seeded batch"}
        stage = {}
        # where the time of a batch goes (device-synchronised wall time of one more batch, stage by stage)
        try:
            stage = stage_breakdown(c4)
        except Exception as ex:          # noqa: BLE001
            stage = {"error": f"{type(ex).__name__}: {ex}"}
        extra["stage_seconds"] = stage
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "c4_steps.json"), "w") as f:
            json.dump({k: counts[k] for k in ("smpl", "object only", "sil", "joint")}, f)
        from vistracker_b200 import default_options, resolve_dims
        extra["c2"] = c2_extra(D, c4.net, c4.sd, c4.dims, min(args.steps, 10)) if D.world == 1 else None
        if D.world == 1:
            try:
                extra["torch_eager_gpu"] = torch_eager_gpu(D, c4.sd, c4.dims)
            except Exception as ex:      # noqa: BLE001
                extra["torch_eager_gpu"] = {"error": f"{type(ex).__name__}: {ex}"}
        if D.world == 1:
            try:
                extra["torch_eager_gpu_c4"] = torch_eager_gpu_c4(D.dev, counts)
            except Exception as ex:      # noqa: BLE001
                extra["torch_eager_gpu_c4"] = {"error": f"{type(ex).__name__}: {ex}"}
            torch.cuda.empty_cache()
        accuracy = c4_accuracy(D) if D.world == 1 else None
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        t0 = time.perf_counter()
        unit = cpu_c4_sample()
        t_cpu = time.perf_counter() - t0
        fps_cpu, per_frame, fps_raster = cpu_c4_fps(unit, counts)
        cpu = {"value": fps_cpu, "unit": UNIT, "cores": cores, "kind": "port", "unit_seconds": unit, "seconds_per_frame": per_frame,
               "value_including_numpy_rasteriser": fps_raster,
               "note": "value excludes the silhouette rasteriser (no CPU implementation exists in the reference; the numpy checker would add sil_step_per_frame per 'sil' frame-step)",
               "sample": (f"EXTRAPOLATED linearly from per-unit times of the oracle restatements on a 4-frame batch ({t_cpu:.0f} s of CPU work: filter 2 frames, "
                          "1 generator projection step on 30000 points, 2 optimize_smpl steps, 2 'object only' + 2 'joint' steps, 1 'sil' frame-step with the numpy "
                          "rasteriser) to the step counts this batch took on the GPU")}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": D.world, "steps": args.steps, "warmup": warmup, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": C4_WORKLOAD, "body_model": BODY_NOTE[BODY], "frames_per_step_per_gpu": c4.frames, "global_frames_per_step": D.world * c4.frames,
                       "sequence_1500_frames": "16 such batches; whole reference batches per rank (SURVEY.md 8(e))",
                       "parallelism": f"frame-batch-parallel x{D.world}, one NCCL all-gather of the [96,169] + [96,13] trajectories per step" if D.world > 1 else "1 GPU",
                       "l2": "working set per step: 0.8 GB of images + 6.8 GB of feature maps for 96 frames, far larger than L2",
                       "generator": "filter_val=10 with the random-init UDF: 2 rounds per target (the minimum a trained network needs)",
                       "step_launch": "one CUDA-graph replay per optimisation step"},
            "roofline": roof, "cpu_baseline": cpu, "accuracy": accuracy,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": c4.h2d_bytes(), "d2h_bytes_per_step": c4.d2h_bytes(),
                    "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps},
            "gpu_launches": launches * args.steps, "clocks": clocks, "extra": extra,
        }
        print(json.dumps(line))
    D.close()


def stage_breakdown(c4: C4):
    """Synchronised wall time of the stages of one more batch (outside the timed region)."""
    import torch
    from vistracker_b200 import recon_driver as RD
    t = {}
    orig = {"generate_all": RD.generate_all, "filter_batch": RD.filter_batch}
    f = c4.fitter
    o_smpl, o_obj = f.optimize_smpl, f.optimize_smpl_object

    def wrap(name, fn):
        def w(*a, **k):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = fn(*a, **k)
            torch.cuda.synchronize(); t[name] = round(time.perf_counter() - t0, 4)
            return r
        return w
    RD.generate_all, RD.filter_batch = wrap("generator", orig["generate_all"]), wrap("filter_batch", orig["filter_batch"])
    f.optimize_smpl, f.optimize_smpl_object = wrap("optimize_smpl", o_smpl), wrap("optimize_smpl_object", o_obj)
    try:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        c4.step(c4.devd, gather=False)
        torch.cuda.synchronize(); t["batch_total"] = round(time.perf_counter() - t0, 4)
    finally:
        RD.generate_all, RD.filter_batch = orig["generate_all"], orig["filter_batch"]
        del f.optimize_smpl, f.optimize_smpl_object
    return t


def run_sequence(D: Dist, args):
    """Strong scaling on a fixed sequence: F frames in batches of 96 sharded with parallel.rank_frames, ONE pass, the trajectories gathered once
    at the end (what the sequence-global stages consume)."""
    import torch
    from vistracker_b200 import parallel
    from vistracker_b200.pipeline import pack_neural
    F_ = args.frames
    mine = parallel.rank_frames(0, F_, C4_FRAMES, D.world, D.rank)
    c4s = {}
    for (s, e) in mine:
        n = e - s
        if n not in c4s:
            c4s[n] = C4(D, frames=n, seed=4 + s)
            c4s[n].step(c4s[n].devd, gather=False)              # warm-up: graph capture, module loading
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sm, ob = [], []
    for (s, e) in mine:
        c = c4s[e - s]
        out = c.step(c.host, gather=False)
        smpl = out["smpl"]
        sm.append(parallel.pack_smplt(torch.cat([smpl.global_pose, smpl.body_pose, smpl.hand_pose], 1).detach(),
                                      torch.cat([smpl.top_betas, smpl.other_betas], 1).detach(), smpl.trans.detach()))
        ob.append(pack_neural(out["obj_R"], out["obj_t"], out["pc_generated"]["object"]["visibility"].to(D.dev)))
    empty = lambda w: torch.zeros(0, w, device=D.dev)
    traj = parallel.gather_trajectory(torch.cat(sm) if sm else empty(169))
    traj_o = parallel.gather_trajectory(torch.cat(ob) if ob else empty(13))
    e1.record()
    D.barrier()
    ms = e0.elapsed_time(e1)
    per_rank = [ms]
    if D.world > 1:
        t = torch.tensor([ms], device=D.dev)
        allt = [torch.zeros_like(t) for _ in range(D.world)]
        D.dist.all_gather(allt, t)
        per_rank = [float(x.item()) for x in allt]
    assert traj.shape == (F_, 169) and traj_o.shape == (F_, 13)
    if D.rank == 0:
        print(json.dumps({"metric": METRIC, "value": F_ / (max(per_rank) * 1e-3), "unit": UNIT, "n_gpus": D.world, "steps": 1, "warmup": 1,
                          "ms_per_step": max(per_rank), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": f"C4 sequence: {F_} frames = {len(parallel.batch_bounds(0, F_, C4_FRAMES))} batches of <= 96, whole batches per rank, "
                                                 "inputs from pinned host memory, one NCCL all-gather of the [T,169] and [T,13] trajectories at the end",
                                     "batches_per_rank": [len(parallel.rank_frames(0, F_, C4_FRAMES, D.world, r)) for r in range(D.world)]},
                          "e2e": {"value": F_ / (max(per_rank) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": sum(c.h2d_bytes() for c in c4s.values()),
                                  "d2h_bytes_per_step": sum(c.d2h_bytes() for c in c4s.values())},
                          "per_rank_ms": per_rank, "gathered": [list(traj.shape), list(traj_o.shape)]}))
    D.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c4", "c2", "c4seq"])
    ap.add_argument("--frames", type=int, default=1536, help="sequence length of --workload c4seq")
    a = ap.parse_args()
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import _inputs as tools_inputs          # noqa: E402  (asset loader shared with the tools; no oracle code)
    sys.modules["tools_inputs"] = tools_inputs
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
