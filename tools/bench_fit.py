"""Timings of the fitting / generation loops at BASELINE config-4 batch shapes (96 frames per batch; 16-frame generator
mini-batches) on one GPU: per-step wall time (synchronised) of optimize_smpl and optimize_smpl_object phases, and the neural
reconstruction (generator) per frame.  Prints one JSON object.

    python tools/bench_fit.py > gpurun_out/fit_loops.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _inputs import load_assets  # noqa: E402
from vistracker_b200 import CHORETriplaneVisibility, default_options, resolve_dims  # noqa: E402
from vistracker_b200.generator import GeneratorTriplaneVis  # noqa: E402
from vistracker_b200.recon_fit import Priors, ReconFitterTriVisFull, SMPLParams  # noqa: E402
from vistracker_b200.render import SilLossROI  # noqa: E402
from vistracker_b200.smpl import LandmarkRegressor, SMPL_Layer  # noqa: E402
from vistracker_b200.synth import synthetic_frames, synthetic_state_dict  # noqa: E402
from vistracker_b200.synth_smpl import synthetic_motion, synthetic_smplh  # noqa: E402

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
a, reg = load_assets()
dims = resolve_dims(default_options())
net = CHORETriplaneVisibility(default_options(), device=dev).eval()
net.load_state_dict(synthetic_state_dict(dims, seed=0))
net.defer_checks = True
model = synthetic_smplh(seed=3)
layer = SMPL_Layer.from_buffers(model, model["parents"], dev)
body25 = LandmarkRegressor(np.stack([reg[0], reg[1]]), reg[2], reg[3], dev)
fitter = ReconFitterTriVisFull(net, Priors(a, dev), torch.from_numpy(a["part_labels"].astype(np.int64)))
res = {"batch_frames": B}


def sync_time(fn, n):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n


# ---- filter on the whole batch (recon_fit_triplane.py:59-60), in sub-batches of 16 to bound activation memory
images, _, crop, _ = synthetic_frames(16, size=512, seed=5, n_points=4, jitter=True)
images = images.repeat(B // 16, 1, 1, 1)[:B]
crop = crop.repeat(B // 16, 1)[:B].to(dev)
t0 = time.perf_counter()
maps = []
for i in range(0, B, 16):
    net.filter(images[i:i + 16].to(dev))
    maps.append(net._maps)
torch.cuda.synchronize()
res["filter_s_per_batch"] = time.perf_counter() - t0
n16 = 16
net._maps = (torch.cat([m[0] for m in maps]), torch.cat([m[1] for m in maps]),
             torch.cat([torch.cat([m[2][v * n16:(v + 1) * n16] for m in maps]) for v in range(3)]),
             torch.cat([torch.cat([m[3][v * n16:(v + 1) * n16] for m in maps]) for v in range(3)]))
del maps

pose, betas, trans = synthetic_motion(B, seed=31)
smpl = SMPLParams(layer, body25, pose, betas, trans)
with torch.no_grad():
    J = smpl.get_landmarks()[0]
body_center = J[:, 8].detach().clone()
qd = {"crop_center": crop, "body_center": body_center}
rng = np.random.Generator(np.random.PCG64(1))
dd = {"part_labels": torch.from_numpy(a["part_labels"].astype(np.int64)).to(dev)[None].repeat(B, 1), "query_dict": qd,
      "pose_init": (pose[:, 3:72] + 0.05 * torch.randn(B, 69)).to(dev), "body_kpts": torch.rand(B, 25, 3, device=dev) * 512}
w = fitter.get_loss_weights()
opt = torch.optim.Adam([smpl.trans, smpl.global_pose, smpl.body_pose, smpl.top_betas, smpl.other_betas], 0.006)


def smpl_step():
    opt.zero_grad()
    loss = fitter.sum_dict(fitter.forward_smpl(smpl, dd, "kpts"), w, 1)
    loss.backward(); opt.step()


res["optimize_smpl_ms_per_step"] = 1e3 * sync_time(smpl_step, 10)

# ---- object phases
from scipy.spatial import ConvexHull  # noqa: E402
p = rng.standard_normal((3000, 3)); p /= np.linalg.norm(p, axis=1, keepdims=True); p *= np.array([0.3, 0.25, 0.2])
tmpl_v = p[:520].astype(np.float32)
faces = ConvexHull(tmpl_v).simplices                      # ~1000 faces, as the BEHAVE templates (recon/opt_utils.py:38-59)
obj_R = torch.eye(3, device=dev)[None].repeat(B, 1, 1).clone().requires_grad_(True)
obj_t = (body_center + torch.tensor([0.3, 0.0, 0.1], device=dev)).clone().requires_grad_(True)
K = SilLossROI.compute_K_roi((424.0, 168.0, 1200.0, 1200.0))[None].repeat(B, 1, 1)
ref = torch.zeros(B, 256, 256); ref[:, 80:176, 96:160] = 1
sil = SilLossROI(torch.ones(B, 256, 256), ref, K, tmpl_v, faces, rend_size=256, device=dev)
od = {"objects": torch.from_numpy(p.astype(np.float32)).to(dev)[None].repeat(B, 1, 1), "query_dict": qd,
      "occ_ratios": torch.rand(B, device=dev) * 0.7 + 0.3, "smpl_center": body_center, "silhouette": sil,
      "trans_init": obj_t.detach().clone(), "obj_s": torch.ones(B, device=dev)}
oopt = torch.optim.Adam([obj_R, obj_t], lr=0.006)
for phase in ("object only", "sil", "joint"):
    def step():
        oopt.zero_grad()
        ld = fitter.forward_step(smpl, od, obj_R, obj_t, od["obj_s"], phase)
        fitter.sum_dict(ld, w, 1).backward(); oopt.step()
    res[f"optimize_object_ms_per_step[{phase}]"] = 1e3 * sync_time(step, 10)
res["faces"] = int(faces.shape[0])

# ---- neural reconstruction: one 16-frame mini-batch, 10 projection steps x (1 seeding round + 1 collecting round) per target
gen = GeneratorTriplaneVis(net, threshold=2.0, filter_val=10.0)      # random-init UDF: accept every in-front point as surface
net._maps = tuple(t[:16] if i < 2 else torch.cat([t[v * B:v * B + 16] for v in range(3)]) for i, t in enumerate(net._maps))
batch = {"crop_center": crop[:16], "body_center": body_center[:16]}
torch.manual_seed(0)
init = gen.get_grid_samples(30000, batch_size=16, body_center=body_center[:16])
torch.cuda.synchronize(); t0 = time.perf_counter()
for tgt in ("human", "object"):
    gen.gen_pc_batch(tgt, init, 4000, batch, num_steps=10)
torch.cuda.synchronize()
res["generator_s_per_16_frames"] = time.perf_counter() - t0
res["generator_note"] = "2 targets x 2 rounds x 10 fused projection steps on 30k / 20k points per frame + resampling glue"
print(json.dumps(res, indent=1))
