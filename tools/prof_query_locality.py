"""Does the ORDER of the query points matter for the fused query-loss launch?  Same 96 x 6890 points (a random cloud around the body centre, as
the synthetic SMPL-H model of the bench produces: no spatial locality along the vertex index) in their native order, and permuted by the
Morton code of the first frame's points (one permutation for all frames, as a template-based ordering would give).
    python tools/prof_query_locality.py [B]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vistracker_b200 import CHORETriplaneVisibility, default_options, resolve_dims  # noqa: E402
from vistracker_b200.recon_driver import filter_batch  # noqa: E402
from vistracker_b200.synth import synthetic_recon_batch, synthetic_state_dict  # noqa: E402
from vistracker_b200.synth_smpl import synthetic_body_mesh  # noqa: E402

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
net = CHORETriplaneVisibility(default_options(), device=dev).eval()
net.load_state_dict(synthetic_state_dict(resolve_dims(default_options()), seed=0))
net.defer_checks = True
h = synthetic_recon_batch(B, seed=4)
with torch.no_grad():
    filter_batch(net, h["images"], chunk=16)
cc, bc = h["crop_center"].to(dev), h["body_center"].to(dev)
rng = np.random.default_rng(0)
cloud = torch.from_numpy((rng.standard_normal((6890, 3)) * np.array([0.25, 0.45, 0.12])).astype(np.float32))
mesh = torch.from_numpy(synthetic_body_mesh()[0]) * 0.9
labels = torch.randint(0, 14, (B, 6890), device=dev)
f = lambda *s: torch.empty(*s, device=dev)


def morton(p):
    q = ((p - p.min(0).values) / (p.max(0).values - p.min(0).values + 1e-9) * 1023).long().clamp(0, 1023)
    def spread(x):
        x = (x | (x << 16)) & 0x030000FF; x = (x | (x << 8)) & 0x0300F00F; x = (x | (x << 4)) & 0x030C30C3; x = (x | (x << 2)) & 0x09249249
        return x
    return torch.argsort(spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2))


def run(tag, pts0, perm=None):
    pts = (pts0[None] + h["body_center"][:, None]).to(dev)
    if perm is not None:
        pts = pts[:, perm.to(dev)]
    pts = pts.contiguous()
    w = torch.tensor([30.0 ** 2, 0.05 ** 2], device=dev)
    vd, vc, g = f(B, 6890), f(B, 6890), f(B, 6890, 3)
    go = lambda: net.enqueue_query_losses_merged(pts, cc, bc, 0, 0.1, labels, w.data_ptr(), 1.0 / (B * 6890), w.data_ptr() + 4, 1.0 / B, vd, vc, g)
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        go()
    e1.record(); torch.cuda.synchronize()
    print(f"{tag}: {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)


run("ellipsoid mesh, ring order      ", mesh)
run("ellipsoid mesh, random order    ", mesh, torch.randperm(6890, generator=torch.Generator().manual_seed(1)))
run("random cloud, native order      ", cloud)
run("random cloud, Morton order      ", cloud, morton(cloud))
run("ellipsoid mesh, Morton order    ", mesh, morton(mesh))
