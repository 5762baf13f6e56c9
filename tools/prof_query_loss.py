"""Timing + phase trace of the fused query-loss launch (vt_query_losses_tc) on B x 6890 vertices (two heads) and B x 3000 object points (one head).
VT_QUERY_TRACE=1 prints the clock64 stamps of CTA (0,0) (epilogue thread 0 | gather warp 0).

    VT_QUERY_TRACE=1 python tools/prof_query_loss.py [B]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vistracker_b200 import CHORETriplaneVisibility, default_options, resolve_dims  # noqa: E402
from vistracker_b200.recon_driver import filter_batch  # noqa: E402
from vistracker_b200.synth import synthetic_recon_batch, synthetic_state_dict  # noqa: E402
from vistracker_b200.synth_smpl import synthetic_body_mesh  # noqa: E402

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
net = CHORETriplaneVisibility(default_options(), device=dev).eval()
net.load_state_dict(synthetic_state_dict(resolve_dims(default_options()), seed=0))
net.defer_checks = True
h = synthetic_recon_batch(B, seed=4)
with torch.no_grad():
    filter_batch(net, h["images"], chunk=16)
bv, _ = synthetic_body_mesh()
verts = (torch.from_numpy(bv)[None] * 0.9 + h["body_center"][:, None]).to(dev).contiguous()
obj = (h["obj_points"][None] + (h["body_center"] + torch.tensor([0.35, 0.0, 0.1]))[:, None]).to(dev).contiguous()
cc, bc = h["crop_center"].to(dev), h["body_center"].to(dev)
f = lambda *s: torch.empty(*s, device=dev)
trace = os.environ.pop("VT_QUERY_TRACE", None)


def run(pts, labels, tag):
    Bn, N = pts.shape[:2]
    out = (f(Bn, N), f(Bn, N, 3), f(Bn, N) if labels is not None else None, f(Bn, N, 3) if labels is not None else None)
    go = lambda: net.enqueue_query_losses(pts, cc, bc, 0 if labels is not None else 1, 0.1 if labels is not None else 0.8, labels, out[0], out[1], out[2], out[3])
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        go()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tiles = Bn * ((N + 127) // 128)
    print(f"{tag}: {ms:.3f} ms for {Bn} x {N} points = {tiles} tiles, {tiles / 148:.1f} per SM, {ms * 1e3 / (tiles / 148):.1f} us per tile-slot", flush=True)
    if trace:
        os.environ["VT_QUERY_TRACE"] = "1"
        go(); torch.cuda.synchronize()
        os.environ.pop("VT_QUERY_TRACE")


labels = torch.randint(0, 14, (B, 6890), device=dev)
run(verts, labels, "df + parts heads on the SMPL vertices")


def run_merged():
    Bn, N = verts.shape[:2]
    w = torch.tensor([30.0 ** 2, 0.05 ** 2], device=dev)
    vd, vc, g = f(Bn, N), f(Bn, N), f(Bn, N, 3)
    go = lambda: net.enqueue_query_losses_merged(verts, cc, bc, 0, 0.1, labels, w.data_ptr(), 1.0 / (Bn * N), w.data_ptr() + 4, 1.0 / Bn, vd, vc, g)
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        go()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tiles = Bn * ((N + 127) // 128)
    print(f"df + parts heads MERGED on the SMPL vertices: {ms:.3f} ms, {ms * 1e3 / (tiles / 148):.1f} us per tile-slot", flush=True)
    if trace:
        os.environ["VT_QUERY_TRACE"] = "1"
        go(); torch.cuda.synchronize()
        os.environ.pop("VT_QUERY_TRACE")


run_merged()
run(obj, None, "df head on the object points")
