"""Per-kernel timings (CUDA events, median of N, L2 flushed between launches) for the hot-path kernels other than the conv,
with the algorithmic bytes / flops of SURVEY.md section 8(d) -> achieved GB/s or TFLOP/s against MEASURED_PEAKS.json.

    python tools/bench_kernels.py > gpurun_out/kernels.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vistracker_b200 import CHORETriplaneVisibility, default_options, resolve_dims  # noqa: E402
from vistracker_b200.synth import synthetic_frames, synthetic_state_dict  # noqa: E402
from vistracker_b200.synth_smpl import synthetic_motion, synthetic_smplh  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
HBM = peaks["hbm_gbs"]


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


rows = []


def row(name, ms, units, unit_name, bytes_per_unit=None, flop_per_unit=None, note=""):
    r = {"kernel": name, "ms": round(ms, 4), "units": units, "unit": unit_name, f"{unit_name}_per_s": units / (ms * 1e-3)}
    if bytes_per_unit:
        r["algorithmic_GBps"] = units * bytes_per_unit / (ms * 1e-3) / 1e9
        r["hbm_frac_of_measured"] = r["algorithmic_GBps"] / HBM
    if flop_per_unit:
        r["TFLOPs"] = units * flop_per_unit / (ms * 1e-3) / 1e12
    r["note"] = note
    rows.append(r)


# ---------------------------------------------------------------- SIF-Net query (C2 shape: 8 frames x 10k points)
dims = resolve_dims(default_options())
net = CHORETriplaneVisibility(default_options(), device=dev).eval()
net.load_state_dict(synthetic_state_dict(dims, seed=0))
images, points, crop, body = synthetic_frames(8, size=512, seed=2, n_points=10000, jitter=True)
net.defer_checks = True
net.filter(images.to(dev))
pts, cc, bc = points.to(dev), crop.to(dev), body.to(dev)
ms = timeit(lambda: net._query_raw(pts, cc, bc))
row("query_fwd_kernel", ms, 8 * 10000, "points", 9856, 1.117e6, "B=8, N=10k; 9 856 B / point, 1.117 MFLOP / point")
g = torch.randn(8, 29, 10000, device=dev)
ms = timeit(lambda: net._query_backward(pts, cc, bc, g))
row("query_bwd_tc_kernel (all heads)", ms, 8 * 10000, "points", 2 * 9728 + 12 + 116 + 12, 3.35e6, "fwd recompute + bwd to points; one gather pair per head")
ms = timeit(lambda: net._query_backward(pts, cc, bc, g, head_mask=5))
row("query_bwd_tc_kernel (df + parts heads: optimize_smpl)", ms, 8 * 10000, "points", 2 * 9728 + 12 + 116 + 12, 2 * 3.35e6 / 5, "head_mask=5")
net.query_on_cuda_cores = True
ms = timeit(lambda: net._query_backward(pts, cc, bc, g))
row("query_bwd_kernel (fp32 FFMA cross-check, all heads)", ms, 8 * 10000, "points", 2 * 9728 + 12 + 116 + 12, 3.35e6, "")
ms = timeit(lambda: net._query_backward(pts, cc, bc, g, head_mask=5))
row("query_bwd_kernel (fp32 FFMA cross-check, df + parts)", ms, 8 * 10000, "points", 2 * 9728 + 12 + 116 + 12, 2 * 3.35e6 / 5, "")
net.query_on_cuda_cores = False
from vistracker_b200.generator import GeneratorTriplaneVis  # noqa: E402
gen = GeneratorTriplaneVis(net)
qi = {"crop_center": cc, "body_center": bc}
ms = timeit(lambda: gen._project_step(pts, qi, 0, False))
row("query_bwd_tc_kernel (projection step, df head only)", ms, 8 * 10000, "points", 2 * 9728 + 24, None, "one approx_surface step")
ms = timeit(lambda: net.filter(images.to(dev)), iters=5)
row("filter (2 encoders, whole launch plan)", ms, 8, "frames", None, 613.46e9, "B=8 512x512; conv-only 613.46 GFLOP / frame")

# ---------------------------------------------------------------- SMPL-H layer, landmarks, SMPL-T fit step
from vistracker_b200.smpl import LandmarkRegressor, SMPL_Layer  # noqa: E402
from _inputs import load_assets, synthetic_fit_problem  # noqa: E402
a, reg = load_assets()
model = synthetic_smplh(seed=3)
layer = SMPL_Layer.from_buffers(model, model["parents"], dev)
for B in (96, 256, 512):
    pose, betas, trans = (t.to(dev) for t in synthetic_motion(B, seed=1))
    ms = timeit(lambda: layer(pose, betas, trans))
    row(f"smpl_fwd (3 launches) B={B}", ms, B, "frames", 676 + 82680 + 624 + 41.7e6 / B, 33e6, "pose+blend GEMM+skin")
    pr = pose.clone().requires_grad_(True)
    v, j, _, _ = layer(pr, betas, trans)
    gv = torch.randn_like(v)
    ms = timeit(lambda: torch.autograd.grad(v, pr, gv, retain_graph=True))
    row(f"smpl_bwd (3 launches + memsets) B={B}", ms, B, "frames", 2 * 82680 + 41.7e6 / B, 33e6, "skin bwd + split-K GEMM + pose bwd")
body25 = LandmarkRegressor(np.stack([reg[0], reg[1]]), reg[2], reg[3], dev)
verts = torch.randn(256, 6890, 3, device=dev)
ms = timeit(lambda: body25(verts))
row("landmarks_fwd B=256", ms, 256, "frames", 82680 + 8481 * 8, None, "body-25, CSR by landmark")

from vistracker_b200.fit_smplt import SMPLHFitter30fps  # noqa: E402
fitter = SMPLHFitter30fps(layer, body25, a)
_, kpts, pose0, betas0, trans0 = synthetic_fit_problem(256, seed=4)
torch.cuda.synchronize(); t0 = time.perf_counter()
out = fitter.fit_batch(pose0, betas0, trans0, kpts, max_iter=20, early_stop=False)
torch.cuda.synchronize(); t_first = time.perf_counter() - t0
t0 = time.perf_counter()
out = fitter.fit_batch(pose0, betas0, trans0, kpts, max_iter=20, early_stop=False)
torch.cuda.synchronize(); t_fit = time.perf_counter() - t0
row("SMPL-T fit, BASELINE config 3 per GPU share (200 Adam steps x 256 frames, CUDA graph)", t_fit * 1e3, 256, "frames", None, None,
    f"{t_fit / 200 * 1e3:.3f} ms / step, first call incl. graph capture {t_first:.2f} s, final loss {out['losses'][-1]:.3f}")

# ---------------------------------------------------------------- rasteriser, chamfer, SO(3)
from vistracker_b200.render import SilhouetteRenderer, TriplaneNrRenderer  # noqa: E402
from scipy.spatial import ConvexHull  # noqa: E402
rng = np.random.Generator(np.random.PCG64(0))
p = rng.standard_normal((1000, 3)); p /= np.linalg.norm(p, axis=1, keepdims=True); p *= np.array([0.4, 0.3, 0.35])
faces = ConvexHull(p).simplices
Bm = 96
K = torch.tensor([[2.0, 0, 0.5], [0, 2.0, 0.5], [0, 0, 1]])[None].repeat(Bm, 1, 1)
rend = SilhouetteRenderer(faces, 256, K, dev)
vv = (torch.from_numpy(p).float()[None].repeat(Bm, 1, 1) + torch.tensor([0, 0, 2.5])).to(dev).requires_grad_(True)
ms = timeit(lambda: rend(vv))
F2 = faces.shape[0]
row(f"raster fwd silhouettes B={Bm} 256^2 F={F2}x2", ms, Bm, "frames", 2 * F2 * 36 + 256 * 256 * 4, None, "setup + tiled face-index map")
img = rend(vv)
gi = torch.randn_like(img)
ms = timeit(lambda: torch.autograd.grad(img, vv, gi, retain_graph=True))
row(f"raster bwd (NMR pseudo-gradient) B={Bm}", ms, Bm, "frames", 2 * F2 * 36 + 2 * 256 * 256 * 4, None, "")
tri = TriplaneNrRenderer(512, dev)
from vistracker_b200.synth_smpl import synthetic_body_mesh  # noqa: E402
_bv, _bf = synthetic_body_mesh()
vs = torch.from_numpy(_bv)[None].repeat(8, 1, 1) + 0.01 * torch.randn(8, 1, 3)
ms = timeit(lambda: tri.render_3views(_bf, vs), iters=5)
row("triplane occupancy 3 views x 1024^2 (2x SSAA) B=8, 13776x2 faces, body-like closed mesh", ms, 8, "frames", 3 * (27552 * 36 + 512 * 512), None,
    "tiles skip 256-face chunks whose bounding box misses them; VT_RASTER_CULL=0 scans every face per tile")
os.environ["VT_RASTER_CULL"] = "0"
ms = timeit(lambda: tri.render_3views(_bf, vs), iters=5)
os.environ.pop("VT_RASTER_CULL")
row("triplane occupancy, same, exhaustive face scan per tile (VT_RASTER_CULL=0)", ms, 8, "frames", 3 * (27552 * 36 + 512 * 512), None, "")
from vistracker_b200.geom import chamfer_distance_ragged, project_so3  # noqa: E402
xs = [torch.randn(int(n), 3, device=dev) for n in rng.integers(20, 300, 400)]
ys = [torch.randn(int(n), 3, device=dev) for n in rng.integers(20, 300, 400)]
ms = timeit(lambda: chamfer_distance_ragged(xs, ys))
row("ragged chamfer, 400 cloud pairs of 20-300 points (incl. host packing)", ms, 400, "pairs", None, None, "")
M = torch.randn(96, 3, 3, device=dev)
ms = timeit(lambda: project_so3(M))
row("so3 projection B=96", ms, 96, "matrices", 72, None, "latency-bound: one thread per matrix")
# ---------------------------------------------------------------- SmoothNet stage (8(f) N1): 1500-frame trajectory
from vistracker_b200.smooth import SMPLTSmoother, ObjrotSmoother, WINDOW  # noqa: E402
_g = np.load(os.path.join(ROOT, "tests", "golden", "smooth_small.npz"))
_sm = SMPLTSmoother({k[6:]: torch.from_numpy(_g[k]) for k in _g.files if k.startswith("smplt.")}, device=dev)
_T = 1500
_poses, _betas, _trans = torch.randn(_T, 156, device=dev) * 0.3, torch.randn(_T, 10, device=dev), torch.randn(_T, 3, device=dev)
ms = timeit(lambda: _sm.smooth(_poses, _betas, _trans))
_rows = (_T - WINDOW + 1) * 147
row("SmoothNet SMPL-T smoothing, 1500 frames (pack + 2 x clips MLP + window mean + unpack)", ms, _T, "frames", None, _rows * 2 * 81920 / _T,
    f"{_rows} (window, channel) rows x 81.9 kMAC, fp32 FFMA")
# ---------------------------------------------------------------- evaluation Chamfer (8(f) N4): 10 000 samples per mesh
from vistracker_b200.geom import eval_chamfer_distance  # noqa: E402
_xa, _ya = torch.randn(16, 10000, 3, device=dev), torch.randn(16, 10000, 3, device=dev)
ms = timeit(lambda: eval_chamfer_distance(_xa, _ya))
row("evaluation Chamfer, 16 frames x (10 000 vs 10 000 points), both directions", ms, 16, "frames", 2 * 10000 * 12 + 2 * 10000 * 4, 2 * 1e8 * 8,
    "brute force, other cloud streamed through shared memory; 8 flop per pair")
# ---------------------------------------------------------------- HVOP-Net (8(f) N2): autoregressive in-filling of a 1500-frame sequence
from vistracker_b200.infill import CondMotionInfillAutoreg, ConditionalMInfiller, default_infill_options  # noqa: E402
from vistracker_b200.synth import synthetic_infill_sequence, synthetic_infill_state_dict  # noqa: E402
_opt = default_infill_options()
_inf = ConditionalMInfiller(_opt, device=dev).load_state_dict(synthetic_infill_state_dict(_opt, 1))
_seq = synthetic_infill_sequence(_T, seed=3)
_occ = _seq[4]
_seq_d = [torch.from_numpy(a).to(dev) for a in _seq[:4]]
_drv = CondMotionInfillAutoreg(_inf)
ms = timeit(lambda: _drv.infill(*_seq_d, _occ))
_clips = len(_drv.clip_plan(_T))
_mac_tok = sum(p.numel() for e in (_inf.enc_smpl, _inf.enc_obj, _inf.enc_joint) for p in e.layers) + 180 * (128 + 32 + 160) * 2 * 0 + 2 * 180 * (2 * 128 + 2 * 32 + 4 * 160)
row(f"HVOP-Net autoregressive in-filling, 1500 frames ({_clips} serial clips x 21 launches, one CUDA graph)", ms, _T, "frames", None,
    _clips * 180 * 2 * _mac_tok / _T, "latency-bound: clip i+1 is seeded by clip i; fp32 FFMA; ms includes the mask upload and output clones")
print(json.dumps({"peaks": peaks, "rows": rows}, indent=1))
