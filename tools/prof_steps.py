"""Per-kernel time inside one optimisation step of the joint fit at BASELINE config-4 shapes (96 frames): every C-ABI call of an eagerly
launched step is bracketed by CUDA events (launch gaps excluded), averaged over N steps.  Prints one JSON object.

    python tools/prof_steps.py [frames] > gpurun_out/prof_steps.json
"""
import collections
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _inputs import load_assets  # noqa: E402
from vistracker_b200 import CHORETriplaneVisibility, _lib, default_options, resolve_dims  # noqa: E402
from vistracker_b200.recon_driver import filter_batch  # noqa: E402
from vistracker_b200.recon_fit import Priors, ReconFitterTriVisFull, SMPLParams  # noqa: E402
from vistracker_b200.recon_steps import ObjectFitStep, SmplRefineStep  # noqa: E402
from vistracker_b200.render import SilLossROI  # noqa: E402
from vistracker_b200.smpl import LandmarkRegressor, SMPL_Layer  # noqa: E402
from vistracker_b200.synth import synthetic_recon_batch, synthetic_state_dict  # noqa: E402
from vistracker_b200.synth_smpl import synthetic_smplh, synthetic_smplh_surface  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 96
N = 10
dev = torch.device("cuda", 0)
a, reg = load_assets()
net = CHORETriplaneVisibility(default_options(), device=dev).eval()
net.load_state_dict(synthetic_state_dict(resolve_dims(default_options()), seed=0))
net.defer_checks = True
model = synthetic_smplh(seed=3) if os.environ.get("VT_BENCH_BODY", "surface") == "cloud" else synthetic_smplh_surface(seed=3)   # as bench.py
layer = SMPL_Layer.from_buffers(model, model["parents"], dev)
body25 = LandmarkRegressor(np.stack([reg[0], reg[1]]), reg[2], reg[3], dev)
h = synthetic_recon_batch(B, seed=4)
fitter = ReconFitterTriVisFull(net, Priors(a, dev), torch.from_numpy(a["part_labels"].astype(np.int64)), scan=(h["obj_verts"].numpy(), h["obj_faces"].numpy()))
c = lambda k: h[k].to(dev)
with torch.no_grad():
    filter_batch(net, h["images"], chunk=16)
qd = {"crop_center": c("crop_center"), "body_center": c("body_center")}
smpl = SMPLParams(layer, body25, h["pose"], h["betas"], h["trans"])
dd = {"part_labels": fitter.part_labels.to(dev)[None].repeat(B, 1), "query_dict": qd, "body_kpts": c("body_kpts"), "pose_init": c("pose")[:, 3:72].clone()}

spans = collections.defaultdict(list)
orig_call = _lib.call


def traced(name, *args):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); orig_call(name, *args); e.record()
    spans[name].append((s, e))


def measure(tag, fn, res):
    fn(); torch.cuda.synchronize()                # warm-up
    spans.clear()
    _lib.call = traced
    try:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(N):
            fn()
        e1.record(); torch.cuda.synchronize()
    finally:
        _lib.call = orig_call
    per = {k: round(sum(s.elapsed_time(e) for s, e in v) / N, 4) for k, v in spans.items()}
    res[tag] = {"ms_per_step_eager_incl_gaps": round(e0.elapsed_time(e1) / N, 4), "kernel_ms_sum": round(sum(per.values()), 4), "per_call_ms": dict(sorted(per.items(), key=lambda kv: -kv[1]))}


res = {"frames": B}
st = SmplRefineStep(fitter, smpl, dd, 64)
st._upload_schedule([[fitter.LOSS_WEIGHTS[k] for k in ("df_h", "pose", "hand", "part", "pinit", "j2d", "stemp")] + [0.0] * 9 + [0.006, 0, 1.0, 0.001] + [0.0] * 12])
st._start(); st._set_row(0)
measure("optimize_smpl step", st._enqueue_step, res)
# graph replay of the same step
st._run(0, st._enqueue_step, True); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(N):
    st._run(0, st._enqueue_step, True)
e1.record(); torch.cuda.synchronize()
res["optimize_smpl step"]["ms_per_step_graph"] = round(e0.elapsed_time(e1) / N, 4)

obj_t = (c("body_center") + torch.tensor([0.35, 0.0, 0.1], device=dev)).clone().requires_grad_(True)
od = {"smpl": smpl, "query_dict": qd, "obj_R": c("obj_rot_init").clone().requires_grad_(True), "obj_t": obj_t, "obj_s": torch.ones(B, device=dev),
      "objects": c("obj_points")[None].repeat(B, 1, 1).contiguous(), "occ_ratios": c("occ_ratios"), "images": c("images")}
sil = fitter._silhouette(od)
with torch.no_grad():
    verts = smpl()[0].detach()
ost = ObjectFitStep(fitter, verts, od, 64, inject_noise=False, seed=1)
W = fitter.LOSS_WEIGHTS
row = lambda on, lr0, lr1, ph, k: [W[n] if n in on else 0.0 for n in ("otemp", "ovtemp", "mask", "scale", "trans", "object", "contact")] + [0.0] * 9 + [lr0, lr1, ph, 1e-4, 0.0, k, 1e-40] + [0.0] * 9
ost._upload_schedule([row(("otemp", "ovtemp", "scale", "object"), 0.002, 0.006, 0.0, 1.0), row(("otemp", "ovtemp", "mask", "scale", "trans"), 0.006, 0.006, 1.0, 1.0),
                      row(("otemp", "ovtemp", "scale", "object", "contact"), 0.0, 0.002, 2.0, 10.0)])
ost._start()
for ph, name in enumerate(("object only", "sil", "joint")):
    ost._set_row(ph)
    if ph == 2:
        ost.enqueue_pose()
        pairs = fitter._first_joint_contacts(od, ost.buf["object"], verts)
        ost.set_contact_pairs(pairs)
        res["contact"] = None if pairs is None else {"human_rows": int(pairs[0].numel()), "object_rows": int(pairs[1].numel()), "clouds": int(pairs[2].numel() - 1)}
    measure(f"object step [{name}]", lambda ph=ph: ost._enqueue_step(ph), res)
    ost.step(ph); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(N):
        ost.step(ph)
    e1.record(); torch.cuda.synchronize()
    res[f"object step [{name}]"]["ms_per_step_graph"] = round(e0.elapsed_time(e1) / N, 4)
res["sil"] = {"template_faces": int(h["obj_faces"].shape[0]), "rend_size": sil.renderer.image_size}
print(json.dumps(res, indent=1))
