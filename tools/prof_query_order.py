"""Does the ORDER of the query points matter for the fused fitting-loss launch (vt_query_losses_tc)?  optimize_smpl queries the 6890 SMPL
vertices of every frame at every Adam step; the kernel is gather-bound, so points that are neighbours in the launch should be neighbours
in the feature maps.  Times forward + backward of ``query_losses`` on B x 6890 points for: random order, mesh (ring) order, and a 3-D
Morton order of the same points.

    python tools/prof_query_order.py [B]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vistracker_b200 import CHORETriplaneVisibility, default_options, resolve_dims  # noqa: E402
from vistracker_b200.synth import synthetic_frames, synthetic_state_dict  # noqa: E402
from vistracker_b200.synth_smpl import synthetic_body_mesh  # noqa: E402

def morton_order(points: torch.Tensor, bits: int = 10) -> torch.Tensor:
    """[B, N, 3] -> [B, N] int64 permutation that sorts every frame's points along a 3-D Z-order curve (``bits`` per axis over the frame's
    bounding box).  A launch-order hint for the gather-bound query kernels: points that are adjacent in the launch then read adjacent
    texels of all four feature-map projections (perspective image plane and the three orthographic planes).  Values are unaffected."""
    p = points.detach().float()
    lo, hi = p.amin(1, keepdim=True), p.amax(1, keepdim=True)
    q = ((p - lo) / (hi - lo).clamp_min(1e-12) * ((1 << bits) - 1)).long().clamp_(0, (1 << bits) - 1)
    code = torch.zeros(p.shape[:2], dtype=torch.int64, device=p.device)
    for b in range(bits):
        for a in range(3):
            code |= ((q[..., a] >> b) & 1) << (3 * b + a)
    return torch.argsort(code, dim=1)


dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dims = resolve_dims(default_options())
net = CHORETriplaneVisibility(default_options(), device=dev).eval()
net.load_state_dict(synthetic_state_dict(dims, seed=0))
net.defer_checks = True
images, _, crop, body = synthetic_frames(16, size=512, seed=5, n_points=4, jitter=True)
images = images.repeat(B // 16, 1, 1, 1)[:B]
crop, body = crop.repeat(B // 16, 1)[:B].to(dev), body.repeat(B // 16, 1)[:B].to(dev)
net.filter(images.to(dev))
bv, _ = synthetic_body_mesh()
verts = torch.from_numpy(bv).to(dev)[None] + body[:, None, :]                                   # [B, 6890, 3] in the camera frame
labels = torch.randint(0, 14, (B, 6890), device=dev)
g = torch.Generator(device=dev).manual_seed(0)


def run(pts, lab):
    p = pts.clone().requires_grad_(True)
    vd, vc = net.query_losses(p, crop_center=crop, df_channel=0, clamp_max=0.1, part_labels=lab, body_center=body)
    (vd.mean() + vc.sum(-1).mean()).backward()
    return p.grad


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


perm_rand = torch.stack([torch.randperm(6890, device=dev, generator=g) for _ in range(B)])
perm_mort = morton_order(verts)
take = lambda x, perm: torch.gather(x, 1, perm[..., None].expand(-1, -1, x.shape[-1])) if x.dim() == 3 else torch.gather(x, 1, perm)
ref = run(verts, labels)
for name, perm in (("mesh (ring) order", None), ("random order", perm_rand), ("3-D Morton order", perm_mort)):
    pts, lab = (verts, labels) if perm is None else (take(verts, perm), take(labels, perm))
    ms = timeit(lambda: run(pts, lab))
    gr = run(pts, lab)
    if perm is not None:                                                                        # same gradients, permuted
        back = torch.empty_like(gr).scatter_(1, perm[..., None].expand(-1, -1, 3), gr)
        err = float((back - ref).abs().max() / ref.abs().max())
    else:
        err = 0.0
    print(f"{name:20s}: {ms:7.3f} ms per forward+backward of {B} x 6890 points ({ms * 1e3 / B:.1f} us / frame), grad mismatch vs mesh order {err:.1e}")
