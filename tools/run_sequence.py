"""The data flow of scripts/demo.sh on ONE synthetic sequence, every stage on the device, shortened optimisation loops: an integration run of
the drop-in objects chained the way the reference's scripts chain them through files (not a benchmark: iteration counts are cut).

  1 SMPL-T pre-fit                    fit_smplt.SMPLHFitter30fps
  2 SmoothNet + re-fit                smooth.SMPLTSmoother, fit_smplt.SMPLHFitterSmoothed
  3 triplane rendering                render.TriplaneNrRenderer
  - frame preparation                 frameio.prepare_images
  4 SIF-Net, neural reconstruction    recon_driver.fit_recon_batch(neural_only=True)  -> parallel.gather_trajectory of the [T,13] block
  5 object SmoothNet + HVOP-Net       pipeline.object_rotation_stage
  6 joint optimisation                recon_driver.fit_recon_batch
  7 outputs + evaluation              io.save_* / pack_recon, evaluate.evaluate_sequence

    python tools/run_sequence.py [frames=80] [outdir]
"""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _inputs import load_assets, synthetic_fit_problem  # noqa: E402
from vistracker_b200 import CHORETriplaneVisibility, default_options, io as vio, parallel, resolve_dims  # noqa: E402
from vistracker_b200.evaluate import evaluate_sequence  # noqa: E402
from vistracker_b200.fit_smplt import SMPLHFitter30fps, SMPLHFitterSmoothed  # noqa: E402
from vistracker_b200.frameio import prepare_images  # noqa: E402
from vistracker_b200.generator import GeneratorTriplaneVis  # noqa: E402
from vistracker_b200.infill import CondMotionInfillAutoreg, ConditionalMInfiller, default_infill_options  # noqa: E402
from vistracker_b200.pipeline import object_rotation_stage, pack_neural  # noqa: E402
from vistracker_b200.recon_driver import fit_recon_batch, scale_body_kpts  # noqa: E402
from vistracker_b200.recon_fit import Priors, ReconFitterTriVisFull, SMPLParams, smplh_pose  # noqa: E402
from vistracker_b200.render import SilLossROI, TriplaneNrRenderer  # noqa: E402
from vistracker_b200.smooth import ObjrotSmoother, SMPLTSmoother  # noqa: E402
from vistracker_b200.smpl import LandmarkRegressor, SMPL_Layer  # noqa: E402
from vistracker_b200.synth import synthetic_camera_frame, synthetic_infill_state_dict, synthetic_state_dict  # noqa: E402
from vistracker_b200.synth_smpl import synthetic_body_mesh  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 80
outdir = sys.argv[2] if len(sys.argv) > 2 else tempfile.mkdtemp(prefix="vt_seq_")
dev = torch.device("cuda", 0)
BS = 16                                   # frames per SIF-Net / joint-optimisation batch in this run
times = {}


def stage(name, t0):
    torch.cuda.synchronize()
    times[name] = round(time.perf_counter() - t0, 3)
    print(f"[{name}] {times[name]:.3f} s", flush=True)


# ---------------------------------------------------------------- models and synthetic inputs
a, reg = load_assets()
model, kpts, pose0, betas0, trans0 = synthetic_fit_problem(T, seed=7)
layer = SMPL_Layer.from_buffers(model, model["parents"], dev)
body25 = LandmarkRegressor(np.stack([reg[0], reg[1]]), reg[2], reg[3], dev)
priors = Priors(a, dev)
gs = np.load(os.path.join(ROOT, "tests", "golden", "smooth_small.npz"))
sd_smplt = {k[6:]: torch.from_numpy(gs[k]) for k in gs.files if k.startswith("smplt.")}
sd_objrot = {k[7:]: torch.from_numpy(gs[k]) for k in gs.files if k.startswith("objrot.")}
net = CHORETriplaneVisibility(default_options(), device=dev).eval()
net.load_state_dict(synthetic_state_dict(resolve_dims(default_options()), seed=0))
net.defer_checks = True
iopt = default_infill_options()
hvop = ConditionalMInfiller(iopt, device=dev).load_state_dict(synthetic_infill_state_dict(iopt, seed=1))
frames = [f"t{i:04d}.000" for i in range(T)]
image_paths = [os.path.join("/data", "Date03_Sub03_chairwood_synth", f, "k1.color.jpg") for f in frames]

# ---------------------------------------------------------------- 1: SMPL-T pre-fit
t0 = time.perf_counter()
fit1 = SMPLHFitter30fps(layer, body25, a).fit_batch(pose0, betas0, trans0, kpts, max_iter=3, early_stop=False)
stage("1 smplt fit (30 steps)", t0)
vio.save_smplt_fits([os.path.join(outdir, "seq", f, "k1.smplfit_temporal.pkl") for f in frames], fit1["pose"], fit1["betas"], fit1["trans"])

# ---------------------------------------------------------------- 2: SmoothNet on the gathered trajectory + re-fit
t0 = time.perf_counter()
traj = parallel.gather_trajectory(parallel.pack_smplt(fit1["pose"], fit1["betas"], fit1["trans"]))
sm = SMPLTSmoother(sd_smplt, device=dev).smooth_trajectory(traj)
pose_s = smplh_pose(sm["poses"], priors.hand_mean).to(dev)
fit2 = SMPLHFitterSmoothed(layer, body25, a).fit_batch(pose_s, sm["betas"], sm["trans"], kpts, max_iter=2, early_stop=False)
stage("2 smoothnet + refit (20 steps)", t0)
vio.pack_smplt(os.path.join(outdir, "recon_smplt-smoothed", "seq_k1.pkl"), frames, "male", fit2["pose"], fit2["betas"], fit2["trans"])

# ---------------------------------------------------------------- 3: triplane rendering of the smoothed SMPL-T
t0 = time.perf_counter()
with torch.no_grad():
    verts, jtr, _, _ = layer(fit2["pose"], th_betas=fit2["betas"], th_trans=fit2["trans"])
    body_center = body25(verts)[:, 8].contiguous()
_, body_faces = synthetic_body_mesh()                     # connectivity with SMPL's counts (the synthetic model's own faces are random)
tri = TriplaneNrRenderer(512, dev)
masks = torch.cat([tri.render_3views(body_faces, verts[s:s + BS] - body_center[s:s + BS, None]) for s in range(0, T, BS)])
stage("3 triplane rendering", t0)
vio.save_triplane_png([os.path.join(outdir, "seq", f, "k1.smooth_triplane.png") for f in frames[:2]], masks[:2])

# ---------------------------------------------------------------- frame preparation (decoded uint8 frames -> network input)
cam = [synthetic_camera_frame(1536, 2048, seed=s) for s in range(2)]
rgb, person, obj = (torch.from_numpy(np.stack([c[k] for c in cam])).to(dev) for k in range(3))
t0 = time.perf_counter()
images, crop_center = [], []
for s in range(0, T, BS):
    n = min(BS, T - s)
    idx = torch.arange(n, device=dev) % 2
    im, cc = prepare_images(rgb[idx], person[idx], obj[idx], (masks[s:s + n].permute(0, 2, 3, 1) * 255).contiguous())
    images.append(im); crop_center.append(cc)
images, crop_center = torch.cat(images), torch.cat(crop_center)
stage("- frame preparation", t0)

# ---------------------------------------------------------------- 4: SIF-Net neural reconstruction per batch, gathered for the sequence stages
fitter = ReconFitterTriVisFull(net, priors, torch.from_numpy(a["part_labels"].astype(np.int64)))
gen = GeneratorTriplaneVis(net, threshold=2.0, filter_val=10.0)      # random-init UDF: accept every in-front point as surface
bv, _ = synthetic_body_mesh(rings=20, segments=20, radii=(0.3, 0.25, 0.2))
obj_points = torch.from_numpy(bv)
pca_init = torch.linalg.qr(torch.randn(3, 3, generator=torch.Generator().manual_seed(1)))[0]
t0 = time.perf_counter()
local = []
for s in range(0, T, BS):
    data = {"images": images[s:s + BS], "crop_center": crop_center[s:s + BS], "body_center": body_center[s:s + BS]}
    pc = fit_recon_batch(fitter, gen, data, None, None, obj_points, neural_only=True)["pc_generated"]
    local.append(pack_neural(pc["object"]["pca_axis"], pc["object"]["centers"][:, 3:], pc["object"]["visibility"]).to(dev))
    if s == 0:
        folders = vio.output_folders(outdir, image_paths[:BS], "test-release")
        vio.save_neural_recon(folders, 1, pc)
neural = parallel.gather_trajectory(torch.cat(local))
stage("4 neural reconstruction", t0)
# the visibility head of a random-init network is meaningless: impose a profile with one occluded span so that HVOP-Net has work to do
neural[:, 12] = 0.8
neural[T // 3: T // 3 + T // 5, 12] = 0.1
vio.pack_recon(os.path.join(outdir, "recon_test-release", "seq_k1.pkl"), frames, "male", "test-release", neural[:, :9], neural[:, 9:12], neural[:, 12:13])

# ---------------------------------------------------------------- 5: object SmoothNet + HVOP-Net
t0 = time.perf_counter()
obj_trans = neural[:, 9:12] + body_center
st5 = object_rotation_stage(neural, fit2["pose"], fit2["trans"], obj_trans, pca_init.to(dev), ObjrotSmoother(sd_objrot, device=dev),
                            CondMotionInfillAutoreg(hvop), occ_thres=0.5)
stage("5 object smoothnet + hvop-net", t0)
print("   in-filled:", st5["infilled"], " det range:", float(torch.linalg.det(st5["obj_angles"].double()).min()), float(torch.linalg.det(st5["obj_angles"].double()).max()))

# ---------------------------------------------------------------- 6: joint optimisation per batch (loops cut to 1 + 1 iterations)
t0 = time.perf_counter()
kp = scale_body_kpts(kpts.to(dev), crop_center)
faces_o = torch.from_numpy(np.asarray(synthetic_body_mesh(rings=20, segments=20)[1]))
res = []
for s in range(0, T, BS):
    n = min(BS, T - s)
    data = {"images": images[s:s + n], "crop_center": crop_center[s:s + n], "body_center": body_center[s:s + n]}
    K = SilLossROI.compute_K_roi((424.0, 168.0, 1200.0, 1200.0))[None].repeat(n, 1, 1)
    ref = torch.zeros(n, 256, 256); ref[:, 80:176, 96:160] = 1
    sil = SilLossROI(torch.ones(n, 256, 256), ref, K, bv, faces_o.numpy(), rend_size=256, device=dev)
    init = lambda human_t, s=s, n=n: SMPLParams(layer, body25, fit2["pose"][s:s + n], fit2["betas"][s:s + n], fit2["trans"][s:s + n])
    out = fit_recon_batch(fitter, gen, data, init, kp[s:s + n], obj_points, obj_rot_init=st5["obj_angles"][s:s + n].transpose(1, 2), silhouette=sil,
                          occ_ratios=neural[s:s + n, 12], max_iter=1, steps_per_iter=1)
    res.append(out)
    folders = vio.output_folders(outdir, image_paths[s:s + n], "test-releasev2")
    vio.save_smpl_params(folders, 1, out["smpl"].pose, out["smpl"].betas, out["smpl"].trans)
    vio.save_object_params(folders, 1, out["obj_R"], out["obj_t"], out["obj_s"])
stage("6 joint optimisation (cut loops)", t0)

# ---------------------------------------------------------------- 7: evaluation of the result against the step-5 initialisation (a sanity number)
t0 = time.perf_counter()
with torch.no_grad():
    sv = torch.cat([r["smpl"]()[0] for r in res])
    ov = torch.cat([ReconFitterTriVisFull.transform_obj_verts(obj_points.to(dev)[None].repeat(r["obj_R"].shape[0], 1, 1), r["obj_R"], r["obj_t"], r["obj_s"]) for r in res])
    ov0 = ReconFitterTriVisFull.transform_obj_verts(obj_points.to(dev)[None].repeat(T, 1, 1), st5["obj_angles"].transpose(1, 2), obj_trans, torch.ones(T, device=dev))
    errs, kept, _ = evaluate_sequence(sv, ov, verts, ov0, window=30, sample_num=None)
stage("7 evaluation", t0)
print(json.dumps({"frames": T, "outdir": outdir, "stage_seconds": times, "files_written": sum(len(f) for _, _, f in os.walk(outdir)),
                  "mean_errors_cm(smpl_cd, obj_cd, smpl_v2v, obj_v2v)": [round(float(x), 3) for x in errs.mean(0)]}))
