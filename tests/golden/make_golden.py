"""Generate the golden vectors in this directory from the UNMODIFIED reference code.

Run in the build container only (the reference is not present on the GPU box):

    cd /root/repo && python tests/golden/make_golden.py [--ref /root/reference]

It imports ``model.CHORETriplaneVisibility`` / ``SMPL_Layer`` straight from the read-only reference tree
(stubbing the unvendored, unused third-party imports exactly as SURVEY.md appendix B describes), loads the
seeded synthetic checkpoint of ``vistracker_b200.synth`` and stores inputs-by-seed + reference outputs:

* ``sifnet_keys.json``      -- the reference ``state_dict()`` key order and shapes (706 entries)
* ``sifnet_small.npz``      -- B=2, 64x64 frames, 301 points: all eight feature maps, all five heads and
                               d(sum clamp(df_h,2))/d(points), d(sum clamp(df_o,2))/d(points)
* ``sifnet_c1.npz``         -- BASELINE config 1 (1 frame 512x512, 2000 points): the five heads in full and
                               every 8th pixel of each feature map
* ``smpl_small.npz``        -- SMPL_Layer.forward on the synthetic SMPL-H model, B=5, outputs + gradients
* ``eval_chamfer.npz``      -- recon/eval/chamfer_distance.py (sklearn kd-tree) on three cloud pairs, all three directions
* ``infill_small.npz``      -- (``--only infill``) ConditionalMInfiller forward on two batches and CondMotionInfillAutoreg.test (the
  reference's own autoregressive clip loop, file IO redirected to a temporary folder) on a 400-frame synthetic sequence.
* ``frameio_small.npz``     -- (``--only frameio``) BehaveDataset.prepare_image_crop / BaseDataset.crop / compose_images on three synthetic
  frames (centred, running past the right/bottom border, past the top/left one) with the real cv2 4.13; direct cv2.resize / bbox vectors.
* ``interp_small.npz``      -- (``--only interp``) BaseInterpolator.compute_missing_inds / interp_slerp / interp_lerp (SLERP baseline).
* ``generator_small.npz``   -- (``--only generator``) the reference's GeneratorTriplaneVis.get_grid_samples / approx_surface / gen_pc_batch
  executed on the CPU (instance created without ``__init__``, which only loads a checkpoint and moves the model to CUDA).
* ``render_views.npz``      -- (``--only render``) TriplaneNrRenderer.transform_view for the three views.
* ``roi_small.npz``         -- (``--only roi``) make_bbox_square, SilLossROI.to_original_bbox / compute_K_roi / cvt_masks.
* ``eval_seq.npz``          -- (``--only evalseq``) VideoPackedEvaluator.eva_seq run on an in-memory synthetic sequence (alignment windows,
  frames without a reconstruction, Chamfer on the vertices, v2v, acceleration errors).
* ``io_formats.npz``        -- (``--only io``) what the reference's writers put on disk (save_neural_recon, save_outputs, save_results).
* ``pack_formats.npz``      -- (``--only pack``) the reference's pack_recon.py / pack_smplt.py run on per-frame files written by this package.
* ``infill_io.npz``         -- (``--only infill_io``) MotionInfillTester.save_output on a small pack.
* ``recon_loop.npz``        -- (``--only reconloop``) ReconFitterBehave.optimize_smpl run unpatched on the CPU, two schedules (per-step terms, totals, final parameters, scale).
* ``recon_obj_loop.npz``    -- (``--only reconobjloop``) ReconFitterTriVisFull.optimize_smpl_object run on the CPU through all three phases.
* ``driver_small.npz``      -- (``--only driver``) ReconFitterBase.scale_body_kpts, ReconFitterBehave.combine_mini_batches.
* ``smooth_small.npz``      -- (``--only smooth``) SmoothNetSMPL / SmoothNet through SMPLTSmoother / ObjrotSmoother pre- and
                               post-processing on a 90-frame synthetic trajectory, window 64, + the rotation conversions
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.dont_write_bytecode = True


def _stub(name, **kw):
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


def import_reference(ref_root: str):
    sk = _stub("skimage"); sk.measure = _stub("skimage.measure")                       # model/mesh_util.py:1
    ch = _stub("chumpy", Ch=object); ch.ch = _stub("chumpy.ch", MatVecMult=None)
    ps = _stub("psbody"); ps.mesh = _stub("psbody.mesh", Mesh=object, MeshViewer=object)
    _stub("psbody.mesh.sphere", Sphere=object)
    os.chdir(ref_root)                       # PATHS.yml / config/ are opened relative to cwd
    sys.path.insert(0, ref_root)


def sifnet_goldens(out_dir: str):
    from config.config_loader import load_configs                                     # reference
    from model import CHORETriplaneVisibility                                         # reference
    from vistracker_b200.config import resolve_dims
    from vistracker_b200.synth import synthetic_frames, synthetic_state_dict

    with contextlib.redirect_stdout(io.StringIO()):
        opt = load_configs("tri-vis-l2")
        net = CHORETriplaneVisibility(opt).eval()
    keys = [(k, list(v.shape)) for k, v in net.state_dict().items()]
    with open(os.path.join(out_dir, "sifnet_keys.json"), "w") as f:
        json.dump(keys, f)
    sd = synthetic_state_dict(resolve_dims(opt), seed=0)
    net.load_state_dict(sd, strict=True)
    for p in net.parameters():
        p.requires_grad = False

    def run(images, points, crop, body, with_grad):
        with torch.no_grad():
            net.filter(images)
        maps = {"im_feat": net.im_feat_list[0], "tmpx": net.tmpx}
        for v in range(3):
            maps[f"tri_feat{v}"] = net.triplane_feat_list[v][0]
            maps[f"tri_tmpx{v}"] = net.triplane_tmpx[v]
        out = {}
        grads = {}
        for name, idx in (("grad_h", 0), ("grad_o", 1)):
            pts = points.clone().requires_grad_(with_grad)
            net.query(pts, crop_center=crop, body_center=body)
            df, pca, parts, centers, vis = net.get_preds()
            if with_grad:
                torch.clamp(df[:, idx], max=2.0).sum().backward()     # recon/gen/generator.py:88-90
                grads[name] = pts.grad.detach().numpy().copy()
        out.update(df=df, pca=pca, parts=parts, centers=centers, vis=vis)
        out = {k: v.detach().numpy().copy() for k, v in out.items()}
        out.update(grads)
        feat, xy = net.query_features(points, crop, body_center=body)
        out["features"] = feat.detach().numpy().copy()
        out["xy"] = xy.detach().numpy().copy()
        return {k: v.detach().numpy().copy() for k, v in maps.items()}, out

    # small case: everything stored
    images, points, crop, body = synthetic_frames(2, size=64, seed=11, n_points=301, jitter=True)
    maps, out = run(images, points, crop, body, True)
    np.savez_compressed(os.path.join(out_dir, "sifnet_small.npz"), **maps, **out)
    print("sifnet_small:", {k: v.shape for k, v in {**maps, **out}.items()})

    # BASELINE config 1: 1 frame 512x512, 2000 points
    images, points, crop, body = synthetic_frames(1, size=512, seed=0, n_points=2000, jitter=False)
    maps, out = run(images, points, crop, body, True)
    maps = {k: np.ascontiguousarray(v[:, :, ::8, ::8]) for k, v in maps.items()}
    out.pop("features")
    np.savez_compressed(os.path.join(out_dir, "sifnet_c1.npz"), **maps, **out)
    print("sifnet_c1:", {k: v.shape for k, v in {**maps, **out}.items()})


def smpl_goldens(out_dir: str):
    from lib_smpl.smplpytorch.smplpytorch.pytorch.smpl_layer import SMPL_Layer       # reference
    from vistracker_b200.synth_smpl import synthetic_smplh, synthetic_motion

    model = synthetic_smplh(seed=3)
    L = SMPL_Layer.__new__(SMPL_Layer)
    torch.nn.Module.__init__(L)
    L.hands, L.num_joints, L.kintree_parents = True, 52, list(model["parents"])
    for k in ("th_betas", "th_shapedirs", "th_posedirs", "th_v_template", "th_J_regressor", "th_weights"):
        L.register_buffer(k, model[k])
    pose, betas, trans = synthetic_motion(5, seed=5)
    pose[0, 3:6] = 0.0            # exercises the theta -> 0 branch of Rodrigues (norm(theta + 1e-8))
    pose.requires_grad_(True); betas.requires_grad_(True); trans.requires_grad_(True)
    verts, jtr, v_posed, naked = L(pose, th_betas=betas, th_trans=trans, th_offsets=torch.zeros(5, 6890, 3))
    rng = np.random.Generator(np.random.PCG64(17))
    gv = torch.from_numpy(rng.standard_normal(tuple(verts.shape), dtype=np.float32))
    gj = torch.from_numpy(rng.standard_normal(tuple(jtr.shape), dtype=np.float32))
    ((verts * gv).sum() + (jtr * gj).sum()).backward()
    np.savez_compressed(os.path.join(out_dir, "smpl_small.npz"),
                        verts=verts.detach().numpy(), jtr=jtr.detach().numpy(), v_posed=v_posed.detach().numpy(),
                        g_pose=pose.grad.numpy(), g_betas=betas.grad.numpy(), g_trans=trans.grad.numpy())
    print("smpl_small: verts", tuple(verts.shape), "jtr", tuple(jtr.shape))


def fit_smplt_goldens(out_dir: str):
    """Runs the reference's own compute_loss / sum_dict / optimiser constructors (unbound, on CPU) inside the loop
    structure of BaseFitter.fit_one_batch (preprocess/fit_SMPLH_kpts.py:131-175) on a 12-frame synthetic problem."""
    for mod in ("behave", "behave.frame_data", "lib_smpl.smpl_generator"):
        _stub(mod, FrameDataReader=object, SMPLHGenerator=object)
    torch.Tensor.cuda = lambda self, *a, **k: self                              # the priors call .cuda() at construction
    from lib_smpl.smplpytorch.smplpytorch.pytorch.smpl_layer import SMPL_Layer       # reference
    from lib_smpl.wrapper_pytorch import SMPLPyTorchWrapperBatchSplitParams          # reference
    from lib_smpl.body_landmark import load_regressors                               # reference
    from preprocess.fit_SMPLH_30fps import SMPLHFitter30fps                          # reference
    import lib_smpl.th_hand_prior as hp_mod                                          # reference
    d = list(hp_mod.HandPrior.__init__.__defaults__)                                  # only the default DEVICE is changed
    hp_mod.HandPrior.__init__.__defaults__ = tuple("cpu" if x == "cuda:0" else x for x in d)

    sys.path.insert(0, os.path.dirname(HERE))
    from fit_problem import synthetic_fit_problem                                    # tests/fit_problem.py
    B = 12
    model, kpts, pose0, betas0, trans0 = synthetic_fit_problem(B, seed=9)
    L = SMPL_Layer.__new__(SMPL_Layer); torch.nn.Module.__init__(L)
    L.hands, L.num_joints, L.kintree_parents = True, 52, list(model["parents"])
    for k in ("th_betas", "th_shapedirs", "th_posedirs", "th_v_template", "th_J_regressor", "th_weights"):
        L.register_buffer(k, model[k])
    S = SMPLPyTorchWrapperBatchSplitParams.__new__(SMPLPyTorchWrapperBatchSplitParams); torch.nn.Module.__init__(S)
    P = torch.nn.Parameter
    S.global_pose, S.body_pose, S.hand_pose = P(pose0[:, :3].clone()), P(pose0[:, 3:66].clone()), P(pose0[:, 66:].clone())
    S.top_betas, S.other_betas, S.trans = P(betas0[:, :2].clone()), P(betas0[:, 2:].clone()), P(trans0.clone())
    S.offsets = P(torch.zeros(B, 6890, 3))
    S.smpl = L
    S.body25_reg_torch, S.face_reg_torch, S.hand_reg_torch = load_regressors("assets", batch_size=B)
    S.verts = S.jtr = S.tposed = S.naked = None
    F = SMPLHFitter30fps.__new__(SMPLHFitter30fps)
    F.fx, F.fy, F.cx, F.cy = 979.7844, 979.840, 1018.952, 779.486
    weights = F.get_loss_weights()
    pose_init = pose0.clone()
    out = {}
    # step-0 loss terms and gradients
    ld = F.compute_loss(S, kpts, pose_init)
    F.sum_dict(ld, weights, 0).backward()
    for k, v in ld.items():
        out[f"loss0_{k}"] = np.float64(v.item())
    out["g0_pose"] = torch.cat([S.global_pose.grad, S.body_pose.grad], 1).numpy().copy()
    out["g0_betas"] = torch.cat([S.top_betas.grad, S.other_betas.grad], 1).numpy().copy()
    out["g0_trans"] = S.trans.grad.numpy().copy()
    # the optimisation loop (no IO, no early stop before it > 30)
    optimizer = F.init_globalpose_optimizer(S)
    losses, record = [], (1, 10, 80, 81, 100)
    step = 0
    for it in range(10):
        if it == F.get_globalopt_iters():
            optimizer = F.init_allpose_optimizer(S)
        for i in range(10):
            optimizer.zero_grad()
            ld = F.compute_loss(S, kpts, pose_init)
            loss = F.sum_dict(ld, weights, it // 3)
            loss.backward(); optimizer.step()
            losses.append(loss.item()); step += 1
            if step in record:
                out[f"pose_{step}"] = torch.cat([S.global_pose, S.body_pose, S.hand_pose], 1).detach().numpy().copy()
                out[f"betas_{step}"] = torch.cat([S.top_betas, S.other_betas], 1).detach().numpy().copy()
                out[f"trans_{step}"] = S.trans.detach().numpy().copy()
    out["losses"] = np.array(losses, np.float64)
    np.savez_compressed(os.path.join(out_dir, "fit_smplt_small.npz"), **out)
    print("fit_smplt_small: losses", losses[0], "->", losses[-1])


def _recon_setup():
    """The UNMODIFIED reference fitter classes on the CPU with a bare ``self``: import stubs for the un-vendored third-party packages, the
    reference SIF-Net with the seeded synthetic checkpoint, a synthetic SMPL-H layer, the seeded problem of tests/recon_problem.py."""
    from argparse import Namespace
    for n in ("trimesh", "igl", "open3d", "zstd", "neural_renderer"):
        _stub(n)
    from oracle.geom_ref import chamfer_ragged

    class Pointclouds:                                   # stand-in for pytorch3d.structures.Pointclouds
        def __init__(self, pts): self.pts = pts
    _stub("pytorch3d"); _stub("pytorch3d.loss", chamfer_distance=lambda a, b: (chamfer_ragged(a.pts, b.pts), None))
    _stub("pytorch3d.structures", Pointclouds=Pointclouds, Meshes=None); _stub("pytorch3d.ops", knn_points=None, sample_points_from_meshes=None)
    _stub("mesh_intersection"); _stub("mesh_intersection.bvh_search_tree", BVH=object); _stub("mesh_intersection.loss")
    _stub("detectron2"); _stub("detectron2.structures", BitMasks=object, BoxMode=object, Boxes=object); _stub("detectron2.structures.boxes", BoxMode=object)
    for m in list(sys.modules):
        if m.startswith(("psbody", "pytorch3d", "detectron2", "mesh_intersection", "skimage", "chumpy")):
            sys.modules[m].__path__ = []
    sys.modules["psbody.mesh"].MeshViewers = object
    torch.Tensor.cuda = lambda self, *a, **k: self
    import lib_smpl.th_hand_prior as hp_mod
    hp_mod.HandPrior.__init__.__defaults__ = tuple("cpu" if x == "cuda:0" else x for x in hp_mod.HandPrior.__init__.__defaults__)
    from config.config_loader import load_configs
    from model import CHORETriplaneVisibility
    from model.camera import KinectColorCamera
    from lib_smpl.smplpytorch.smplpytorch.pytorch.smpl_layer import SMPL_Layer
    from lib_smpl.wrapper_pytorch import SMPLPyTorchWrapperBatchSplitParams
    from lib_smpl.body_landmark import load_regressors
    import recon.recon_fit_trivis_full as M                                                # reference
    import recon.recon_fit_base as MB
    MB.chamfer_distance = M.chamfer_distance = sys.modules["pytorch3d.loss"].chamfer_distance
    MB.Pointclouds = M.Pointclouds = Pointclouds
    from vistracker_b200.config import resolve_dims
    from vistracker_b200.synth import synthetic_state_dict
    sys.path.insert(0, os.path.dirname(HERE))
    from recon_problem import B, make_problem

    d = make_problem()
    with contextlib.redirect_stdout(io.StringIO()):
        opt = load_configs("tri-vis-l2")
        net = CHORETriplaneVisibility(opt).eval()
    net.load_state_dict(synthetic_state_dict(resolve_dims(opt), seed=0))
    for p in net.parameters():
        p.requires_grad = False
    with torch.no_grad():
        net.filter(d["images"])
    L = SMPL_Layer.__new__(SMPL_Layer); torch.nn.Module.__init__(L)
    L.hands, L.num_joints, L.kintree_parents = True, 52, list(d["model"]["parents"])
    for k in ("th_betas", "th_shapedirs", "th_posedirs", "th_v_template", "th_J_regressor", "th_weights"):
        L.register_buffer(k, d["model"][k])

    def make_smpl():
        S = SMPLPyTorchWrapperBatchSplitParams.__new__(SMPLPyTorchWrapperBatchSplitParams); torch.nn.Module.__init__(S)
        P = torch.nn.Parameter
        S.global_pose, S.body_pose, S.hand_pose = P(d["pose"][:, :3].clone()), P(d["pose"][:, 3:66].clone()), P(d["pose"][:, 66:].clone())
        S.top_betas, S.other_betas, S.trans = P(d["betas"][:, :2].clone()), P(d["betas"][:, 2:].clone()), P(d["trans"].clone())
        S.offsets = P(torch.zeros(B, 6890, 3)); S.smpl = L; S.faces = None
        S.body25_reg_torch, S.face_reg_torch, S.hand_reg_torch = load_regressors("assets", batch_size=B)
        S.verts = S.jtr = S.tposed = S.naked = None
        S.betas = torch.cat([S.top_betas, S.other_betas], 1); S.pose = torch.cat([S.global_pose, S.body_pose, S.hand_pose], 1)
        return S

    F = M.ReconFitterTriVisFull.__new__(M.ReconFitterTriVisFull)
    F.camera, F.net_in_size, F.debug, F.z_0, F.device, F.obj_scale = KinectColorCamera(1200), 512, False, 2.2, "cpu", 1.0
    F.part_labels, F.collision_loss, F.args = d["labels"], False, Namespace(model_name="chore-triplane-vis")
    F.part_names = [str(i) for i in range(14)]
    return d, net, make_smpl, F, B


def recon_goldens(out_dir: str):
    """forward_smpl (phase 'kpts') and forward_step (phases 'object only', 'joint') of the UNMODIFIED reference fitter classes,
    called unbound on CPU with a bare ``self`` (recon/recon_fit_behave.py:467-513, recon/recon_fit_trivis_full.py:193-270).
    pytorch3d's chamfer_distance is not installable -> the contact term goes through oracle.geom_ref.chamfer_ragged."""
    d, net, make_smpl, F, B = _recon_setup()
    weights = F.get_loss_weights()
    qd = {"crop_center": d["crop"], "body_center": d["body_center"]}
    out = {}
    # ---- forward_smpl, phase 'kpts'
    S = make_smpl()
    dd = {"part_labels": d["labels"][None].repeat(B, 1), "net": net, "query_dict": qd, "pose_init": d["pose_init"], "body_kpts": d["body_kpts"]}
    ld = F.forward_smpl(S, dd, "kpts")
    F.sum_dict(ld, weights, 2 / 3).backward()
    for k, v in ld.items():
        out[f"smpl_{k}"] = np.float64(v.item())
    out["smpl_g_pose"] = torch.cat([S.global_pose.grad, S.body_pose.grad], 1).numpy().copy()
    out["smpl_g_betas"] = torch.cat([S.top_betas.grad, S.other_betas.grad], 1).numpy().copy()
    out["smpl_g_trans"] = S.trans.grad.numpy().copy()
    # ---- forward_step, phases 'object only' and 'joint'
    for phase in ("object only", "joint"):
        S = make_smpl()
        R_, t_ = d["obj_R"].clone().requires_grad_(True), d["obj_t"].clone().requires_grad_(True)
        dd = {"objects": d["objects"], "query_dict": qd, "occ_ratios": d["occ"], "smpl_center": d["smpl_center"],
              "df_obj_h": d["df_obj_h"], "df_hum_o": d["df_hum_o"], "parts_obj": d["parts_obj"]}
        real_rand = torch.rand
        torch.rand = lambda *a, **k: d["noise"].clone()            # decopose_axis: rot + 1e-4 * torch.rand(B, 3, 3)
        try:
            ld = F.forward_step(net, S, dd, R_, t_, d["obj_s"], phase)
        finally:
            torch.rand = real_rand
        F.sum_dict(ld, weights, 1 if phase == "object only" else 4 / 3).backward()
        tag = "obj" if phase == "object only" else "joint"
        for k, v in ld.items():
            out[f"{tag}_{k}"] = np.float64(v.item())
        out[f"{tag}_g_R"], out[f"{tag}_g_t"] = R_.grad.numpy().copy(), t_.grad.numpy().copy()
    np.savez_compressed(os.path.join(out_dir, "recon_small.npz"), **out)
    print("recon_small:", {k: (float(v) if np.ndim(v) == 0 else v.shape) for k, v in out.items()})


def smooth_goldens(out_dir: str):
    """SmoothNet stage: the reference's SmoothNetSMPL / SmoothNet modules (seeded init, eval mode) driven through the
    unbound SMPLTSmoother / ObjrotSmoother preprocess_input + post_processing and slide_window_to_sequence -> smooth_small.npz."""
    y = _stub("yacs"); y.config = _stub("yacs.config", CfgNode=dict)
    _stub("smoothnet.core.evaluate_config", parse_args=None)                         # needs yacs; only parse_args is imported from it
    b = _stub("behave"); b.frame_data = _stub("behave.frame_data", FrameDataReader=object)
    b.utils = _stub("behave.utils", load_template=None)
    _stub("recon.pca_util", PCAUtil=object)
    from smoothnet.models import SmoothNet, SmoothNetSMPL                             # reference
    from smoothnet.smooth_base import SmootherBase                                    # reference
    from smoothnet.smooth_smplt import SMPLTSmoother                                  # reference
    from smoothnet.smooth_objrot import ObjrotSmoother                                # reference
    import smoothnet.utils.geometry_utils as G                                       # reference

    W = 64

    def shim(cls):
        s = types.SimpleNamespace(slide_window_size=W, slide_window_step=1, device="cpu")
        for name in ("seq2batches", "merge_paths", "smplh2smpl_pose"):
            if hasattr(cls, name):
                setattr(s, name, types.MethodType(getattr(cls, name), s))
        return s

    rng = np.random.default_rng(7)
    L = 90
    t = np.arange(L)[:, None] / 15.0
    poses = (0.4 * np.sin(t * rng.uniform(0.5, 2.0, (1, 156)) + rng.uniform(0, 6, (1, 156))) + 0.05 * rng.standard_normal((L, 156))).astype(np.float32)
    poses[:, :3] += np.array([2.6, 0.3, -0.2], np.float32)                            # a global orientation near pi: the quaternion branches
    poses[5, 3:6] = 0.0                                                               # an exactly-zero joint rotation
    betas = (rng.standard_normal((1, 10)) * 0.5 + 0.01 * rng.standard_normal((L, 10))).astype(np.float32)
    trans = (np.array([[0.1, -0.2, 2.3]]) + 0.3 * np.sin(t * np.array([[0.7, 1.1, 0.4]])) + 0.01 * rng.standard_normal((L, 3))).astype(np.float32)
    frames = np.array([f"t{i:04d}.000" for i in range(L)])

    torch.manual_seed(11)
    net = SmoothNetSMPL(window_size=W, output_size=W, hidden_size=512, res_hidden_size=16, num_blocks=1, dropout=0.5).eval()
    for p in net.parameters():
        p.data.mul_(3.0)                                                              # make the residual path matter at random init
    s = shim(SMPLTSmoother)
    data = SMPLTSmoother.preprocess_input(s, {"poses": poses, "betas": betas, "trans": trans, "frames": frames})
    with torch.no_grad():
        inp = data["input_data"]
        den = net(inp.permute(0, 2, 1)).permute(0, 2, 1)
    rec = SMPLTSmoother.post_processing(s, data, den.clone(), inp.clone())
    out = {"L": L, "W": W, "poses_in": poses, "betas_in": betas, "trans_in": trans, "input_data": inp.numpy(), "denoised_clips": den.numpy(),
           "poses_out": rec["poses"], "betas_out": rec["betas"], "trans_out": rec["trans"]}
    for k, v in net.state_dict().items():
        out["smplt." + k] = v.numpy()

    torch.manual_seed(12)
    onet = SmoothNet(window_size=W, output_size=W, hidden_size=512, res_hidden_size=16, num_blocks=1, dropout=0.5).eval()
    for p in onet.parameters():
        p.data.mul_(3.0)
    aa = (np.array([[0.3, 1.2, -0.4]]) + 0.5 * np.sin(t * np.array([[0.9, 0.6, 1.3]])) + 0.05 * rng.standard_normal((L, 3))).astype(np.float32)
    rot = G.batch_rodrigues(torch.from_numpy(aa)).numpy()                             # "real" rotation matrices [L, 3, 3]
    so = shim(ObjrotSmoother)
    odata = ObjrotSmoother.preprocess_input(so, {"obj_rot": rot, "neural_visibility": np.zeros(L), "frames": frames})
    with torch.no_grad():
        oin = odata["input_data"]
        oden = onet(oin.permute(0, 2, 1)).permute(0, 2, 1)
    orec = ObjrotSmoother.post_processing(so, odata, oden.clone(), oin.clone())
    out.update({"obj_rot_in": rot, "obj_input_data": oin.numpy(), "obj_denoised_clips": oden.numpy(), "obj_angles_out": orec["obj_angles"]})
    for k, v in onet.state_dict().items():
        out["objrot." + k] = v.numpy()
    # conversion functions on their own (branch coverage of rotation_matrix_to_quaternion)
    aa_t = torch.from_numpy(np.concatenate([rng.uniform(-3.1, 3.1, (200, 3)), np.zeros((1, 3)), [[np.pi, 0, 0], [0, np.pi - 1e-3, 0], [0, 0, 3.0]]]).astype(np.float32))
    r6 = G.axis_to_rot6D(aa_t).reshape(-1, 6)
    out.update({"conv_axis": aa_t.numpy(), "conv_rot6d": r6.numpy(), "conv_axis_back": G.rot6D_to_axis(r6.clone()).numpy(),
                "conv_np_rot6d": G.numpy_axis_to_rot6D(aa_t.numpy()).reshape(-1, 6)})
    np.savez_compressed(os.path.join(out_dir, "smooth_small.npz"), **out)
    print("smooth_small.npz:", {k: getattr(v, "shape", v) for k, v in out.items() if not k.startswith(("smplt.", "objrot."))})


def eval_goldens(out_dir: str):
    """Evaluation Chamfer: recon/eval/chamfer_distance.py run as is (sklearn kd-tree) -> eval_chamfer.npz."""
    from recon.eval.chamfer_distance import chamfer_distance                          # reference
    rng = np.random.default_rng(21)
    out = {}
    for i, (n, m) in enumerate(((700, 900), (1, 50), (1500, 1500))):
        x = rng.standard_normal((n, 3)).astype(np.float32) * 0.4
        y = (x[rng.integers(0, n, m)] + 0.02 * rng.standard_normal((m, 3))).astype(np.float32) if i != 1 else rng.standard_normal((m, 3)).astype(np.float32)
        out[f"x{i}"], out[f"y{i}"] = x, y
        for d in ("bi", "x_to_y", "y_to_x"):
            out[f"cd{i}_{d}"] = np.float64(chamfer_distance(x, y, direction=d))
    # Procrustes alignment: compute_transform / compute_similarity_transform of recon/eval/pose_utils.py (psbody is only imported there)
    _stub("psbody"); _stub("psbody.mesh", Mesh=object)
    from recon.eval.pose_utils import compute_similarity_transform, compute_transform    # reference
    for i, n in enumerate((500, 9466)):
        src = rng.standard_normal((n, 3)) * np.array([0.3, 0.8, 0.2]) + np.array([0.1, -0.2, 2.4])
        ang = rng.standard_normal(3); ang *= (0.5 + i) / np.linalg.norm(ang)
        from scipy.spatial.transform import Rotation
        Rt = Rotation.from_rotvec(ang).as_matrix()
        dst = (1.0 + 0.2 * i) * src.dot(Rt.T) + np.array([0.4, 0.1, -0.3]) + 0.01 * rng.standard_normal((n, 3))
        if i == 1:
            dst[:, 0] *= -1.0                                                          # a reflection: the determinant fix must kick in
        R, t, sc, _ = compute_transform(src.astype(np.float32).astype(np.float64), dst.astype(np.float32).astype(np.float64))
        out[f"pa_src{i}"], out[f"pa_dst{i}"] = src.astype(np.float32), dst.astype(np.float32)
        out[f"pa_R{i}"], out[f"pa_t{i}"], out[f"pa_s{i}"] = R, t[:, 0], np.float64(sc)
        out[f"pa_hat{i}"] = compute_similarity_transform(src.astype(np.float32).astype(np.float64), dst.astype(np.float32).astype(np.float64))
    np.savez_compressed(os.path.join(out_dir, "eval_chamfer.npz"), **out)
    print("eval_chamfer.npz:", {k: float(v) for k, v in out.items() if k.startswith("cd")})


def asset_fixtures(out_dir: str, ref_root: str):
    """Numeric assets the reference ships for this path, re-serialised without scipy / pickle:
    the body-25 landmark regressor (COO), the pose / hand priors and the 14-part vertex labels."""
    import pickle
    import warnings
    warnings.filterwarnings("ignore")
    ld = lambda rel: pickle.load(open(os.path.join(ref_root, "assets", rel), "rb"), encoding="latin1")
    out = {}
    for name in ("body25", "face", "hand"):
        m = ld(f"{name}_regressor.pkl").tocoo()
        out[f"{name}_row"], out[f"{name}_col"] = m.row.astype(np.int32), m.col.astype(np.int32)
        out[f"{name}_val"], out[f"{name}_shape"] = m.data.astype(np.float32), np.array(m.shape, np.int32)
    for name in ("body_prior", "lh_prior", "rh_prior"):
        d = ld(f"priors/{name}.pkl")
        out[f"{name}_mean"], out[f"{name}_precision"] = np.asarray(d["mean"], np.float64), np.asarray(d["precision"], np.float64)
    parts = ld("smpl_parts_dense.pkl")
    labels = np.full(6890, -1, np.int8)
    for i, k in enumerate(sorted(parts.keys())):          # dict order == sorted order; label n = n-th key (recon/recon_fit_base.py:315-325)
        labels[np.asarray(parts[k])] = i
    out["part_labels"] = labels
    np.savez_compressed(os.path.join(out_dir, "assets.npz"), **out)
    print("assets:", {k: v.shape for k, v in out.items()})

def infill_goldens(out_dir: str):
    """HVOP-Net: the reference's ConditionalMInfiller (eval mode, config/cmf-k4-lrot.json, seeded synthetic checkpoint of
    vistracker_b200/synth.py) on two batches, and the reference's own autoregressive loop -- CondMotionInfillAutoreg.test, file IO redirected
    to a temporary folder -- on a 400-frame synthetic sequence (a full first clip, 8 strided clips, a short last clip) -> infill_small.npz."""
    import tempfile
    from argparse import Namespace
    import joblib
    b = _stub("behave"); b.utils = _stub("behave.utils", load_template=None); b.frame_data = _stub("behave.frame_data", FrameDataReader=object)
    t = _stub("trainer"); t.__path__ = []; _stub("trainer.train_utils", load_checkpoint=None)     # trainer/ imports trimesh; only load_checkpoint is used
    _stub("lib_smpl", get_smpl=None)
    from recon.pca_util import PCAUtil                                                # reference
    from config.config_loader import load_configs                                     # reference
    from interp.test_cinfill_autoreg import CondMotionInfillAutoreg                   # reference
    from model import ConditionalMInfiller                                            # reference
    from vistracker_b200.synth import synthetic_infill_sequence, synthetic_infill_state_dict

    with contextlib.redirect_stdout(io.StringIO()):
        opt = load_configs("cmf-k4-lrot")
    sd = synthetic_infill_state_dict(opt, seed=21)
    net = ConditionalMInfiller(opt).eval()
    missing = net.load_state_dict(sd, strict=True)
    out = {"opt_json": json.dumps({k: v for k, v in vars(opt).items() if isinstance(v, (int, float, str, bool, list))})}

    rng = np.random.default_rng(3)
    for tag, B, T in (("a", 2, 180), ("b", 1, 47)):
        ds = rng.standard_normal((B, T, opt.dim_smpl)).astype(np.float32)
        do = rng.standard_normal((B, T, opt.dim_obj)).astype(np.float32)
        mo = rng.random((B, T)) < 0.4
        mo[:, :5] = False
        ms = np.zeros((B, T), bool)
        if tag == "b":
            ms[:, 10:20] = True                                                        # the SMPL branch's mask is honoured too
        do = do * (1 - mo[..., None].astype(np.float32))
        with torch.no_grad():
            pred = net(torch.from_numpy(ds), torch.from_numpy(ms), torch.from_numpy(do), torch.from_numpy(mo))
        out.update({f"{tag}_data_smpl": ds, f"{tag}_mask_smpl": ms, f"{tag}_data_obj": do, f"{tag}_mask_obj": mo, f"{tag}_pred": pred.numpy()})

    # the glue between SIF-Net's PCA-axis prediction and the rotation inputs of this stage (test_infiller.py:172-183, smooth_objrot.py:46-57)
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    src = (q * np.array([[1.0], [0.6], [0.3]])).astype(np.float32)                    # rows = scaled, orthogonal template axes
    tgt = np.stack([np.linalg.qr(rng.standard_normal((3, 3)))[0] @ src + 0.05 * rng.standard_normal((3, 3)) for _ in range(40)]).astype(np.float32)
    tgt[3] *= -1                                                                      # a reflected prediction: the det fix of project_so3
    tmpl = (rng.standard_normal((300, 3)) * np.array([0.5, 0.2, 0.35])) @ np.linalg.qr(rng.standard_normal((3, 3)))[0] + np.array([0.1, -0.2, 0.05])
    out.update({"pca_template": tmpl, "pca_components_installed_sklearn": PCAUtil.compute_pca(tmpl)})
    out.update({"pca_src": src, "pca_tgt": tgt,
                "pca_R": PCAUtil.init_object_orientation(torch.from_numpy(tgt), torch.stack([torch.from_numpy(src)] * 40, 0)).numpy()})

    L = 400
    rot6d_smpl, trans_smpl, rot6d_obj, trans_obj, occ = synthetic_infill_sequence(L, seed=5)
    tmp = tempfile.mkdtemp(prefix="vt_infill_")
    seq = "Date03_Sub03_chairwood_synthetic"

    class Tester(CondMotionInfillAutoreg):
        def __init__(self):
            self.device, self.outdir, self.model, self.icap_kid, self.exp_name = "cpu", tmp, net, 2, "cmf-k4-lrot"

        def get_test_files(self, args, recon_name):
            return [os.path.join(tmp, "in.pkl")], [seq]

        def prepare_rot6d(self, args, dat, file, recon_name, seq_name, gt_data):
            return rot6d_obj.copy(), rot6d_smpl.copy()

    joblib.dump({"frames": [f"t{i:04d}.000" for i in range(L)], "trans": trans_smpl.copy(), "obj_trans": trans_obj.copy(),
                 "obj_angles": np.zeros((L, 3, 3), np.float32)}, os.path.join(tmp, "in.pkl"))
    os.makedirs(os.path.join(tmp, "recon_objname"), exist_ok=True)
    joblib.dump({"neural_visibility": np.stack([occ, occ], 1)}, os.path.join(tmp, f"recon_objname/{seq}_k1.pkl"))
    args = Namespace(**vars(opt))
    args.smpl_recon_name, args.obj_recon_name, args.save_name, args.occ_thres, args.occ_pred = "smplname", "objname", "out", 0.5, True
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        Tester().test(args)
    res = joblib.load(os.path.join(tmp, f"recon_out/{seq}_k1.pkl"))
    out.update({"seq_L": L, "seq_rot6d_smpl": rot6d_smpl, "seq_trans_smpl": trans_smpl, "seq_rot6d_obj": rot6d_obj, "seq_trans_obj": trans_obj,
                "seq_occ": occ, "seq_obj_angles": res["obj_angles"], "seq_obj_trans": res["obj_trans"], "seq_obj_scales": res["obj_scales"]})
    np.savez_compressed(os.path.join(out_dir, "infill_small.npz"), **out)
    print("infill_small.npz:", {k: getattr(v, "shape", v) for k, v in out.items() if k != "opt_json"}, missing)

def frameio_goldens(out_dir: str):
    """Test-time frame preparation: the reference's BehaveDataset.prepare_image_crop and BaseDataset.crop /
    compose_images / resize / masks2bbox / center_from_masks (data/train_data.py:143-162, data/base_data.py:139-171,204-265), called unbound on a
    shim, WITH THE REAL OpenCV (cv2 4.13, opencv-python-headless): cv2.resize(INTER_LINEAR) on uint8 and cv2.threshold / findContours /
    boundingRect run as the reference runs them.  Also direct cv2 vectors for the two restated operators: resizes at the production ratio
    (1200 -> 512 = 2.34375, 3 and 1 channels; the full-size result as a SHA-256), a non-square and an identity resize, and the bounding box of
    masks with isolated speckle, a uint8 wrap-around (128 + 128) and soft borders -> frameio_small.npz."""
    import hashlib
    import cv2                                                                         # real
    pkg = _stub("data"); pkg.__path__ = [os.path.join(os.getcwd(), "data")]       # data/__init__.py pulls in trimesh / igl; load the two files only
    ps = _stub("psbody"); ps.mesh = _stub("psbody.mesh", Mesh=object)
    from data.base_data import BaseDataset                                            # reference
    from data.train_data import BehaveDataset                                         # reference
    from vistracker_b200.synth import synthetic_camera_frame

    H, W, CROP, NET = 180, 240, 150, 64                                                # 150 / 64 = 1200 / 512
    out = {"H": H, "W": W, "crop": CROP, "net": NET, "cv2_version": cv2.__version__}
    centers = {"mid": None, "right_bottom": (205.0, 150.0), "top_left": (30.0, 25.0)}
    for i, (tag, ctr) in enumerate(centers.items()):
        rgb, person, obj = synthetic_camera_frame(H, W, seed=30 + i, center=ctr)
        shim = types.SimpleNamespace(CROP_SIZE=np.array([CROP, CROP]), img_size=(NET, NET), dtype=np.float32,
                                     load_masks=lambda f, flip: (person, obj), load_rgb=lambda f, flip: rgb)
        for name in ("crop", "compose_images", "resize", "masks2bbox", "center_from_masks"):
            setattr(shim, name, types.MethodType(getattr(BaseDataset, name), shim))
        images, center = BehaveDataset.prepare_image_crop(shim, "frame.color.jpg", False)
        out[f"{tag}_images"], out[f"{tag}_center"] = images, np.asarray(center)
        out[f"{tag}_crop_rgb"] = BaseDataset.crop(shim, rgb, center, shim.CROP_SIZE)
        bmin, bmax = BaseDataset.masks2bbox(shim, [person, obj])
        out[f"{tag}_bbox"] = np.concatenate([bmin, bmax])
    # ---- cv2.resize vectors (seeded noise + gradients: every interpolation weight pair occurs)
    rng = np.random.Generator(np.random.PCG64(77))
    def noise(h, w, c=None):
        return rng.integers(0, 256, (h, w) if c is None else (h, w, c)).astype(np.uint8)
    for tag, (src, dsize) in {"r300_128_rgb": (noise(300, 300, 3), (128, 128)), "r300_128_mask": (noise(300, 300), (128, 128)),
                              "r300x200_128x96": (noise(200, 300, 3), (128, 96)), "r75_32": (noise(75, 75, 3), (32, 32)),
                              "r64_64": (noise(64, 64, 3), (64, 64)), "r37_100": (noise(37, 37), (100, 100))}.items():
        out[f"{tag}_src"], out[f"{tag}_dst"] = src, cv2.resize(src, dsize, interpolation=cv2.INTER_LINEAR)
    big_rgb, big_person, _ = synthetic_camera_frame(1200, 1200, seed=41)
    out["r1200_512_rgb_sha256"] = hashlib.sha256(cv2.resize(big_rgb, (512, 512), interpolation=cv2.INTER_LINEAR).tobytes()).hexdigest()
    out["r1200_512_mask_sha256"] = hashlib.sha256(cv2.resize(big_person, (512, 512), interpolation=cv2.INTER_LINEAR).tobytes()).hexdigest()
    # ---- masks2bbox vectors
    shim = types.SimpleNamespace()
    boxes = []
    for k in range(4):
        a, b = np.zeros((120, 160), np.uint8), np.zeros((120, 160), np.uint8)
        r = np.random.Generator(np.random.PCG64(90 + k))
        x0, y0 = int(r.integers(5, 60)), int(r.integers(5, 40))
        a[y0:y0 + int(r.integers(8, 50)), x0:x0 + int(r.integers(8, 60))] = 255
        b[y0 + 10:y0 + 10 + int(r.integers(8, 50)), x0 + 20:x0 + 20 + int(r.integers(8, 60))] = int(r.integers(128, 256))
        for _ in range(6):                                                            # isolated speckle, some below the threshold
            yy, xx = int(r.integers(0, 120)), int(r.integers(0, 160))
            a[yy, xx] = int(r.choice([60, 127, 128, 200, 255]))
        if k == 1:                                                                    # 128 + 128 wraps to 0 in uint8 before the clip
            a[100, 150] = 128; b[100, 150] = 128
        out[f"bbox{k}_a"], out[f"bbox{k}_b"] = a, b
        bmin, bmax = BaseDataset.masks2bbox(shim, [a, b])
        boxes.append(np.concatenate([bmin, bmax]))
    out["bbox_ref"] = np.stack(boxes)
    np.savez_compressed(os.path.join(out_dir, "frameio_small.npz"), **out)
    print("frameio_small.npz:", {k: getattr(v, "shape", v) for k, v in out.items()})

def interp_goldens(out_dir: str):
    """SLERP / LERP baseline (interp/interpolate_recon.py): BaseInterpolator's static methods and the quaternion pipeline of interp_seq /
    save_output on a synthetic sequence with three occluded spans (one of them crossing the antipodal hemisphere) -> interp_small.npz."""
    from scipy.spatial.transform import Rotation
    _stub("behave", SCRATCH_PATH="/tmp", GTPACK_PATH="/tmp")
    from interp.interpolate_recon import BaseInterpolator                            # reference
    from vistracker_b200.synth import synthetic_infill_sequence
    L = 220
    _, _, _, trans_obj, occ = synthetic_infill_sequence(L, seed=4, occluded=((20, 45), (90, 140), (170, 181)))
    rng = np.random.default_rng(6)
    t = np.arange(L)[:, None] / 25.0
    rv = np.array([[0.5, 2.6, -0.4]]) + 0.9 * np.sin(t * np.array([[0.8, 0.5, 1.1]])) + 0.02 * rng.standard_normal((L, 3))     # angles near pi: sign flips of q
    R = Rotation.from_rotvec(rv).as_matrix()
    obj_angles = R.transpose(0, 2, 1).copy()
    mask = (occ < 0.3).astype(float)
    end_inds, start_inds = BaseInterpolator.compute_missing_inds(mask)
    rot_q = Rotation.from_matrix(obj_angles.transpose(0, 2, 1)).as_quat()
    frames = [f"t{i:04d}.000" for i in range(L)]
    q = BaseInterpolator.interp_slerp(end_inds, frames, rot_q, start_inds, mute=True)
    tr = BaseInterpolator.interp_lerp(end_inds, frames, trans_obj.astype(np.float64), start_inds, mute=True)
    out = {"obj_angles_in": obj_angles, "occ": occ, "trans_in": trans_obj, "end_inds": end_inds, "start_inds": start_inds, "quat_out": q,
           "obj_angles_out": Rotation.from_quat(q).as_matrix().transpose(0, 2, 1), "trans_out": tr}
    np.savez_compressed(os.path.join(out_dir, "interp_small.npz"), **out)
    print("interp_small.npz:", {k: getattr(v, "shape", v) for k, v in out.items()}, "spans", list(zip(start_inds, end_inds)))

def generator_goldens(out_dir: str):
    """UDF -> point cloud generator: the reference's own GeneratorTriplaneVis methods -- get_grid_samples,
    approx_surface, gen_pc_batch (with parse_preds / compose_outdict) -- executed on the CPU.  ``Generator.__init__`` wants a checkpoint
    directory and a CUDA device (recon/gen/generator.py:28-52), none of which the methods need: the instance is created without it and
    given device='cpu', the thresholds and the reference SIF-Net (seeded synthetic checkpoint).  -> generator_small.npz"""
    t = _stub("trainer"); t.__path__ = []; _stub("trainer.train_utils", convertMillis=None, convertSecs=None, load_checkpoint=None)   # trainer/ imports trimesh
    from config.config_loader import load_configs                                     # reference
    from model import CHORETriplaneVisibility                                         # reference
    from recon.gen.generator_vis import GeneratorTriplaneVis                          # reference
    from vistracker_b200.config import resolve_dims
    from vistracker_b200.synth import synthetic_frames, synthetic_state_dict

    with contextlib.redirect_stdout(io.StringIO()):
        opt = load_configs("tri-vis-l2")
        net = CHORETriplaneVisibility(opt).eval()
    net.load_state_dict(synthetic_state_dict(resolve_dims(opt), seed=0), strict=True)
    for p in net.parameters():
        p.requires_grad = False
    gen = object.__new__(GeneratorTriplaneVis)
    gen.device, gen.threshold, gen.filter_val, gen.sparse_thres, gen.model = torch.device("cpu"), 2.0, 10.0, 0.03, net
    images, _, crop, body = synthetic_frames(2, size=64, seed=11, n_points=4, jitter=True)
    batch = {"images": images, "crop_center": crop, "body_center": body, "path": ["a", "b"]}
    gen.filter(batch)
    torch.manual_seed(123)
    init = gen.get_grid_samples(300, batch_size=2, body_center=body)
    out = {"init": init.numpy().copy()}
    # one call of approx_surface: 3 chained projection steps, both targets
    for name in ("human", "object"):
        s = init.clone().requires_grad_(True)
        surf, preds = gen.approx_surface(net, s, 3, gen.prep_query_input(batch), df_type=name)
        out[f"surf_{name}"] = surf.detach().numpy().copy()
        out[f"surf_df_{name}"] = preds[0].detach().numpy().copy()
    # the whole loop: resampling draws from torch's global CPU generator; filter_val 10 accepts every in-front point of the random-init UDF
    torch.manual_seed(7)
    with contextlib.redirect_stdout(io.StringIO()):
        pc = gen.gen_pc_batch(net, "object", init, 250, batch, num_steps=1, mute=True)
    for k, v in pc.items():
        out[f"pc_{k}"] = v.numpy().copy()
    np.savez_compressed(os.path.join(out_dir, "generator_small.npz"), **out)
    print("generator_small.npz:", {k: v.shape for k, v in out.items()})

def render_goldens(out_dir: str):
    """TriplaneNrRenderer.transform_view (render/render_triplane_nr.py:110-139), the part of the triplane rendering that is the reference's
    own code (the rasteriser behind it, neural_renderer, is not installable) -> render_views.npz."""
    _stub("cv2", setNumThreads=lambda n: None)
    b = _stub("behave"); b.frame_data = _stub("behave.frame_data", FrameDataReader=object); b.kinect_transform = _stub("behave.kinect_transform", KinectTransform=object)
    _stub("lib_smpl.body_landmark", BodyLandmarks=object)
    _stub("neural_renderer", Renderer=object)
    from render.render_triplane_nr import TriplaneNrRenderer                          # reference
    pts = np.random.default_rng(12).standard_normal((40, 3))
    out = {"points": pts}
    for view in ("right", "back", "top"):
        out[view] = TriplaneNrRenderer.transform_view(pts, view)
    out["top_z5"] = TriplaneNrRenderer.transform_view(pts, "top", z_offset=5.0)
    np.savez_compressed(os.path.join(out_dir, "render_views.npz"), **out)
    print("render_views.npz:", {k: v.shape for k, v in out.items()})

def roi_goldens(out_dir: str):
    """The set-up arithmetic of SilLossROI (recon/obj_pose_roi.py:21-75): make_bbox_square (recon/bbox.py:25-46), to_original_bbox (:110-120),
    compute_K_roi (:123-155, torch.cuda.FloatTensor redirected to the CPU type), cvt_masks (:157-170).  The mask crop itself
    (detectron2 BitMasks.crop_and_resize) and cv2's contour bbox are not installable -> not in this golden.  -> roi_small.npz"""
    _stub("cv2", setNumThreads=lambda n: None)
    _stub("neural_renderer", Renderer=object)
    d = _stub("detectron2"); d.structures = _stub("detectron2.structures", BitMasks=object); _stub("detectron2.structures.boxes", BoxMode=object)
    import scipy.ndimage
    if not hasattr(scipy.ndimage, "morphology"):                                   # removed namespace in recent scipy; only an import in the reference
        _stub("scipy.ndimage.morphology", distance_transform_edt=scipy.ndimage.distance_transform_edt)
    _stub("recon.opt_utils", mask2bbox=None)                                         # imports cv2 / psbody at module level; one function is used
    from recon.bbox import make_bbox_square                                           # reference
    from recon.obj_pose_roi import SilLossROI                                         # reference
    torch.cuda.FloatTensor = torch.FloatTensor
    rng = np.random.default_rng(17)
    boxes_xywh = np.stack([rng.uniform(50, 300, 6), rng.uniform(40, 280, 6), rng.uniform(30, 160, 6), rng.uniform(30, 160, 6)], 1)
    squares = make_bbox_square(boxes_xywh.copy(), 0.3)
    centers = rng.uniform(600, 1400, (6, 2))
    orig = np.stack([SilLossROI.to_original_bbox(sq, 1200 / 512, c, 1200) for sq, c in zip(squares, centers)])
    Ks = np.stack([SilLossROI.compute_K_roi(b)[0].numpy() for b in orig])
    K_icap = SilLossROI.compute_K_roi(orig[0], image_width=1920, fx=918.457763671875, fy=918.4373779296875, cx=956.9661865234375, cy=555.944580078125)[0].numpy()
    ps, ob = torch.from_numpy(rng.random((3, 32, 32)).astype(np.float32)), torch.from_numpy(rng.random((3, 32, 32)).astype(np.float32))
    keep = torch.stack([SilLossROI.cvt_masks(None, p, o) for p, o in zip(ps, ob)]).numpy()
    out = {"boxes_xywh": boxes_xywh, "squares": squares, "centers": centers, "orig": orig, "Ks": Ks, "K_icap": K_icap, "ps": ps.numpy(), "ob": ob.numpy(), "keep": keep}
    np.savez_compressed(os.path.join(out_dir, "roi_small.npz"), **out)
    print("roi_small.npz:", {k: v.shape for k, v in out.items()})

def evalseq_goldens(out_dir: str):
    """The evaluation loop itself: the reference's VideoPackedEvaluator.eva_seq (recon/eval/evalvideo_packed.py:29-163)
    executed on a synthetic sequence -- packed-file loading (prep_verts) replaced by in-memory arrays, psbody's Mesh by a two-field stand-in,
    and surface_sampling by the vertices themselves (trimesh is not installed and its samples are unseeded draws anyway) -- so that the
    alignment windows, the handling of frames without a reconstruction, Chamfer (sklearn kd-tree), v2v and the acceleration errors are the
    reference's.  -> eval_seq.npz"""
    from argparse import Namespace

    class SimpleMesh:
        def __init__(self, v=None, f=None):
            self.v, self.f = v, f
    ps = _stub("psbody"); ps.mesh = _stub("psbody.mesh", Mesh=SimpleMesh)
    _stub("trimesh")
    b = _stub("behave"); b.seq_utils = _stub("behave.seq_utils", SeqInfo=lambda seq: Namespace(get_obj_name=lambda: "chairwood"))
    tmpl = SimpleMesh(np.zeros((1, 3)), np.zeros((1, 3), int))
    b.utils = _stub("behave.utils", load_template=lambda name, **kw: tmpl)
    _stub("lib_smpl", SMPL_Layer=lambda **kw: Namespace(th_faces=torch.zeros(1, 3, dtype=torch.long)))
    _stub("recon.recon_data", ReconDataReader=object)
    _stub("recon.opt_utils")
    from recon.eval.evalvideo_packed import VideoPackedEvaluator                      # reference

    rng = np.random.default_rng(33)
    L, Vs, Vo, W = 23, 60, 25, 7
    t = np.arange(L)[:, None, None] / 10.0
    sv_gt = rng.standard_normal((1, Vs, 3)) * np.array([0.3, 0.8, 0.2]) + np.array([0.0, 0.0, 2.3]) + 0.1 * np.sin(t * np.array([1.0, 2.0, 0.5]))
    ov_gt = rng.standard_normal((1, Vo, 3)) * 0.2 + np.array([0.5, 0.1, 2.1]) + 0.1 * np.cos(t * np.array([0.7, 1.3, 0.9]))
    from scipy.spatial.transform import Rotation
    Rm = Rotation.from_rotvec([0.2, -0.4, 0.1]).as_matrix()
    sv_rc = 1.1 * sv_gt.dot(Rm.T) + np.array([0.3, -0.1, 0.2]) + 0.01 * rng.standard_normal((L, Vs, 3))
    ov_rc = 1.1 * ov_gt.dot(Rm.T) + np.array([0.3, -0.1, 0.2]) + 0.03 * rng.standard_normal((L, Vo, 3))
    exist = np.ones(L, bool); exist[[3, 7, 8, 9, 10, 11, 12, 13]] = False               # one whole window (7..13) has no reconstruction
    out = {"sv_gt": sv_gt, "ov_gt": ov_gt, "sv_rc": sv_rc, "ov_rc": ov_rc, "exist": exist, "window": W}

    class Ev(VideoPackedEvaluator):
        def __init__(self):
            self.errors_dict, self.sample_num, self.unit_cvt = {}, 10000, 100

        def prep_verts(self, save_name, seq_name, smplh_layer, temp, tid):
            return {"recon_exist": self._exist, "frames": [f"t{i:04d}.000" for i in range(L)]}, ov_gt, ov_rc, sv_gt, sv_rc

        def surface_sampling(self, m):
            return m.v

    for tag, ex in (("all", np.ones(L, bool)), ("gaps", exist)):
        ev = Ev(); ev._exist = ex
        with contextlib.redirect_stdout(io.StringIO()):
            ev.eva_seq("/data/Date03_Sub03_chairwood_synth", "name", 1, args=Namespace(window=W))
        out[f"errors_{tag}"] = np.asarray(ev.errors_dict["Date03_Sub03_chairwood_synth"])
    # (window <= 0, "no alignment", divides by zero in the packed evaluator's acceleration bookkeeping, :148 -- not a usable mode there)
    np.savez_compressed(os.path.join(out_dir, "eval_seq.npz"), **out)
    print("eval_seq.npz:", {k: getattr(v, "shape", v) for k, v in out.items()})

def io_goldens(out_dir: str):
    """On-disk formats: the reference's own writers executed into a temporary folder and read back --
    ReconFitterBase.save_neural_recon / save_outputs / get_output_paths (recon/recon_fit_base.py:278-313, 830-845) with
    opt_utils.save_smplfits (recon/opt_utils.py:113-141), BaseFitter.save_results (preprocess/fit_SMPLH_kpts.py:228-261) and the
    neural-only joblib pack of preprocess/pack_recon.py:118-133 -> io_formats.npz (file names, key order, dtypes, shapes, values)."""
    import pickle
    import tempfile
    from argparse import Namespace
    for n in ("trimesh", "igl", "open3d", "zstd", "neural_renderer"):
        _stub(n)

    class Pointclouds:
        def __init__(self, pts): self.pts = pts
    _stub("pytorch3d"); _stub("pytorch3d.loss", chamfer_distance=None)
    _stub("pytorch3d.structures", Pointclouds=Pointclouds, Meshes=None); _stub("pytorch3d.ops", knn_points=None, sample_points_from_meshes=None)
    _stub("mesh_intersection"); _stub("mesh_intersection.bvh_search_tree", BVH=object); _stub("mesh_intersection.loss")
    _stub("detectron2"); _stub("detectron2.structures", BitMasks=object, BoxMode=object, Boxes=object); _stub("detectron2.structures.boxes", BoxMode=object)
    for m in list(sys.modules):
        if m.startswith(("psbody", "pytorch3d", "detectron2", "mesh_intersection", "skimage", "chumpy")):
            sys.modules[m].__path__ = []
    written = []

    class RecMesh:                                       # records what psbody's Mesh(...).write_ply would be asked to write
        def __init__(self, v=None, f=None): self.v, self.f = v, f
        def write_ply(self, path): written.append(path)
    sys.modules["psbody.mesh"].Mesh = RecMesh
    sys.modules["psbody.mesh"].MeshViewers = object
    import recon.recon_fit_base as MB                                                  # reference
    import preprocess.fit_SMPLH_kpts as FK                                             # reference
    FK.Mesh = RecMesh
    rng = np.random.default_rng(41)
    B, n = 3, 50
    tmp = tempfile.mkdtemp(prefix="vt_io_")
    paths = [os.path.join("/data/behave", "Date03_Sub03_chairwood_hand", f"t{i:04d}.{(37 * i) % 1000:03d}", "k1.color.jpg") for i in range(B)]
    tf = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32))
    recon_batch = {t: {"points": tf(B, n, 3), "pca_axis": tf(B, 3, 3), "parts": torch.from_numpy(rng.integers(0, 14, (B, n))),
                       "centers": torch.cat([torch.full((B, 3), float("nan")), tf(B, 3)], 1), "visibility": torch.rand(B, 1)} for t in ("human", "object")}
    shim = Namespace(outpath=os.path.join(tmp, "recon"))
    MB.ReconFitterBase.save_neural_recon(shim, paths, recon_batch, "test-release", 1)
    out = {"paths": np.array(paths)}
    for i in range(B):
        f = os.path.join(shim.outpath, "Date03_Sub03_chairwood_hand", os.path.basename(os.path.dirname(paths[i])), "test-release", "k1_densepc.npz")
        d = np.load(f, allow_pickle=True)
        assert sorted(d.files) == ["human", "object"]
        for t in d.files:
            for k, v in d[t].item().items():
                out[f"densepc{i}.{t}.{k}"] = v
        out[f"densepc{i}.order"] = np.array([f"{t}.{k}" for t in d.files for k in d[t].item()])
    for t in ("human", "object"):
        for k, v in recon_batch[t].items():
            out[f"in.{t}.{k}"] = v.numpy()
    # save_outputs: SMPL parameter pickle (+ score) and the object pickle with the rotation re-projected without noise
    pose, betas, trans = tf(B, 156), tf(B, 10), tf(B, 3)
    smpl = Namespace(pose=pose, betas=betas, trans=trans, faces=torch.zeros(1, 3, dtype=torch.long))
    smpl_call = lambda: (torch.zeros(B, 4, 3), None, None, None)
    class SmplShim:
        def __init__(self): self.pose, self.betas, self.trans, self.faces = pose, betas, trans, torch.zeros(1, 3, dtype=torch.long)
        def __call__(self): return smpl_call()
    obj_R = tf(B, 3, 3) * 0.1 + torch.eye(3)
    obj_t, obj_s = tf(B, 3), torch.ones(B)
    fit = Namespace(outpath=shim.outpath, scan=Namespace(v=rng.standard_normal((20, 3)), f=np.zeros((1, 3), int)), device="cpu")
    for name in ("get_output_paths", "transform_object", "transform_obj_verts"):
        setattr(fit, name, types.MethodType(getattr(MB.ReconFitterBase, name), fit))
    fit.decopose_axis = MB.ReconFitterBase.decopose_axis
    MB.ReconFitterBase.save_outputs(fit, SmplShim(), obj_R, obj_t, paths, "test-releasev2", 1, obj_s)
    for i in range(B):
        folder = os.path.join(shim.outpath, "Date03_Sub03_chairwood_hand", os.path.basename(os.path.dirname(paths[i])), "test-releasev2")
        sm, ob = pickle.load(open(os.path.join(folder, "k1.smpl.pkl"), "rb")), pickle.load(open(os.path.join(folder, "k1.object.pkl"), "rb"))
        out[f"smpl{i}.order"], out[f"object{i}.order"] = np.array(list(sm)), np.array(list(ob))
        for k, v in sm.items():
            out[f"smpl{i}.{k}"] = np.asarray(v)
        for k, v in ob.items():
            out[f"object{i}.{k}"] = np.asarray(v)
    out.update({"in.pose": pose.numpy(), "in.betas": betas.numpy(), "in.trans": trans.numpy(), "in.obj_R": obj_R.numpy(), "in.obj_t": obj_t.numpy(), "in.obj_s": obj_s.numpy()})
    # BaseFitter.save_results: per-frame SMPL-T pickle, frames without confident key points are skipped
    scores = torch.rand(B, 25); scores[1] = 0.0
    files = [os.path.join(tmp, "seq", os.path.basename(os.path.dirname(p)), "k1.color.jpg") for p in paths]
    for f in files:
        os.makedirs(os.path.dirname(f), exist_ok=True)
    fk = Namespace(get_outfile=lambda folder, kid: os.path.join(folder, f"k{kid}.smplfit_temporal.pkl"))
    fk.skip_frame = types.MethodType(FK.BaseFitter.skip_frame, fk); fk.save_smpl_mesh = types.MethodType(FK.BaseFitter.save_smpl_mesh, fk)
    with contextlib.redirect_stdout(io.StringIO()):
        FK.BaseFitter.save_results(fk, SmplShim(), "seq", 1, 0, B, scores, files)
    out["smplt.written"] = np.array([os.path.isfile(os.path.join(os.path.dirname(f), "k1.smplfit_temporal.pkl")) for f in files])
    d0 = pickle.load(open(os.path.join(os.path.dirname(files[0]), "k1.smplfit_temporal.pkl"), "rb"))
    out["smplt.order"] = np.array(list(d0))
    for k, v in d0.items():
        out[f"smplt0.{k}"] = np.asarray(v)
    out["in.scores"] = scores.numpy()
    out["ply_requests"] = np.array([os.path.basename(p) for p in written])
    np.savez_compressed(os.path.join(out_dir, "io_formats.npz"), **out)
    print("io_formats.npz:", len(out), "entries;", "ply requests:", sorted(set(out["ply_requests"].tolist())))

def pack_goldens(out_dir: str):
    """Interoperability of the per-frame files with the reference's packers: the files are written by THIS package's
    vistracker_b200.io (k1_densepc.npz, k1.smpl.pkl, k1.object.pkl, k1.smplfit_smoothed.pkl) and then read by the reference's own
    preprocess/pack_recon.py:main (neural-only and full) through its ReconDataReader, and preprocess/pack_smplt.py:main.  Only the frame
    enumeration (behave.frame_data.FrameDataReader, un-vendored) and the SMPL model files behind get_root_joint are stand-ins.
    The packs they write are stored -> pack_formats.npz, which vistracker_b200.io.pack_recon / pack_smplt must reproduce."""
    import tempfile
    from argparse import Namespace
    import joblib
    from vistracker_b200 import io as vio
    tmp = tempfile.mkdtemp(prefix="vt_pack_")
    T = 5
    frames = [f"t{i:04d}.{(41 * i) % 1000:03d}" for i in range(T)]
    seq = "Date03_Sub03_chairwood_hand"

    class FrameDataReader:                                 # behave.frame_data (the BEHAVE toolkit is not vendored by the reference)
        def __init__(self, seq_folder, check_image=False, ext="jpg"):
            self.seq_path, self.seq_name, self.frames = seq_folder, os.path.basename(seq_folder), list(frames)
            self.seq_info = Namespace(get_gender=lambda: "male")
        def __len__(self): return len(self.frames)
        def get_frame_folder(self, idx): return os.path.join(self.seq_path, self.frames[idx])
        def frame_time(self, idx): return self.frames[idx]
    b = _stub("behave"); b.frame_data = _stub("behave.frame_data", FrameDataReader=FrameDataReader)
    _stub("cv2", setNumThreads=lambda n: None)
    root_of = lambda pose, betas, trans: (trans + 0.125).reshape(-1, 1, 3)            # stand-in for SMPL_Layer.get_root_joint (licensed model files)
    layer = Namespace(get_root_joint=root_of)
    _stub("lib_smpl", SMPL_Layer=lambda **kw: layer, get_smpl=lambda *a, **k: layer)
    import preprocess.pack_recon as PR                                                # reference
    import preprocess.pack_smplt as PS                                                # reference

    rng = np.random.default_rng(51)
    f32 = lambda *s: rng.standard_normal(s).astype(np.float32)
    recon_root = os.path.join(tmp, "recon")
    paths = [os.path.join("/data", seq, f, "k1.color.jpg") for f in frames]
    pcs = {t: {"points": torch.from_numpy(f32(T, 30, 3)), "pca_axis": torch.from_numpy(f32(T, 3, 3)), "parts": torch.from_numpy(rng.integers(0, 14, (T, 30))),
               "centers": torch.cat([torch.full((T, 3), float("nan")), torch.from_numpy(f32(T, 3))], 1), "visibility": torch.rand(T, 1)} for t in ("human", "object")}
    pose, betas, trans = f32(T, 156), f32(T, 10), f32(T, 3)
    from scipy.spatial.transform import Rotation
    rot = Rotation.from_rotvec(rng.standard_normal((T, 3))).as_matrix().astype(np.float32)
    obj_t, obj_s = f32(T, 3), np.ones(T, np.float32)
    for name in ("test-release", "test-releasev2"):
        folders = vio.output_folders(recon_root, paths, name)
        vio.save_neural_recon(folders, 1, pcs)
    vio.save_smpl_params(folders, 1, pose, betas, trans)
    vio.save_object_params(folders, 1, rot, obj_t, obj_s)
    out = {"frames": np.array(frames), "in.pca": pcs["object"]["pca_axis"].numpy(), "in.centers": pcs["object"]["centers"].numpy(),
           "in.vis": pcs["object"]["visibility"].numpy(), "in.pose": pose, "in.betas": betas, "in.trans": trans, "in.rot": rot, "in.obj_t": obj_t, "in.obj_s": obj_s}

    def store(tag, d):
        out[f"{tag}.order"] = np.array(list(d))
        for k, v in d.items():
            if isinstance(v, list) and len(v) and not isinstance(v[0], str):
                out[f"{tag}.{k}"], out[f"{tag}.{k}.islist"] = np.stack([np.asarray(x) for x in v], 0), np.array(True)
            else:
                out[f"{tag}.{k}"] = np.asarray(v)

    seq_folder = os.path.join(tmp, "data", seq)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        PR.main(Namespace(seq_folder=seq_folder, out=os.path.join(tmp, "packed"), recon_path=recon_root, save_name="test-release", neural_only=True, test_ids=[1]))
        PR.main(Namespace(seq_folder=seq_folder, out=os.path.join(tmp, "packed"), recon_path=recon_root, save_name="test-releasev2", neural_only=False, test_ids=[1]))
    store("neural", joblib.load(os.path.join(tmp, "packed", "recon_test-release", f"{seq}_k1.pkl")))
    store("full", joblib.load(os.path.join(tmp, "packed", "recon_test-releasev2", f"{seq}_k1.pkl")))
    # pack_smplt reads k1.smplfit_smoothed.pkl from the SEQUENCE folders
    vio.save_smplt_fits([os.path.join(seq_folder, f, "k1.smplfit_smoothed.pkl") for f in frames], pose, betas, trans)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        PS.main(Namespace(seq_folder=seq_folder, mesh_type="smoothed", out=os.path.join(tmp, "packed_smplt"), test_id=1))
    store("smplt", joblib.load(os.path.join(tmp, "packed_smplt", f"{seq}_k1.pkl")))
    np.savez_compressed(os.path.join(out_dir, "pack_formats.npz"), **out)
    print("pack_formats.npz:", {k: out[k].tolist() for k in out if k.endswith(".order")})

def infill_io_goldens(out_dir: str):
    """MotionInfillTester.save_output (interp/test_infiller.py:129-143) run unbound on a small pack (both branches: in-filled rotations,
    and save_orig when HVOP-Net skipped the sequence) -> infill_io.npz."""
    import tempfile
    from argparse import Namespace
    import joblib
    b = _stub("behave"); b.utils = _stub("behave.utils", load_template=None); b.frame_data = _stub("behave.frame_data", FrameDataReader=object)
    t = _stub("trainer"); t.__path__ = []; _stub("trainer.train_utils", load_checkpoint=None)
    _stub("lib_smpl", get_smpl=None)
    from interp.test_infiller import MotionInfillTester                               # reference
    tmp = tempfile.mkdtemp(prefix="vt_infio_")
    rng = np.random.default_rng(61)
    L = 6
    dat = {"poses": rng.standard_normal((L, 156)), "betas": rng.standard_normal((L, 10)), "trans": rng.standard_normal((L, 3)),
           "obj_angles": rng.standard_normal((L, 3, 3)), "obj_trans": rng.standard_normal((L, 3)), "obj_scales": np.zeros(L), "gender": "male",
           "frames": [f"t{i:04d}.000" for i in range(L)]}
    rot = torch.from_numpy(rng.standard_normal((L, 3, 3)).astype(np.float32))
    tr = torch.from_numpy(rng.standard_normal((L, 3)).astype(np.float32))
    out = {"rot_pred": rot.numpy(), "trans_pred": tr.numpy()}
    for k, v in dat.items():
        out[f"in.{k}"] = np.asarray(v)
    shim = Namespace(exp_name="cmf-k4-lrot")
    for tag, kw in (("filled", dict(rot_pred=rot, trans_pred=tr)), ("orig", dict(rot_pred=None, trans_pred=None, save_orig=True))):
        f = os.path.join(tmp, tag, "seq_k1.pkl")
        with contextlib.redirect_stdout(io.StringIO()):
            MotionInfillTester.save_output(shim, {k: (v.copy() if hasattr(v, "copy") else v) for k, v in dat.items()}, f, **kw)
        d = joblib.load(f)
        out[f"{tag}.order"] = np.array(list(d))
        for k, v in d.items():
            out[f"{tag}.{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(out_dir, "infill_io.npz"), **out)
    print("infill_io.npz:", out["filled.order"].tolist())

def _unsplit_container(d, L, B):
    """A ``SMPLPyTorchWrapperBatch`` (lib_smpl/wrapper_pytorch.py:23-70) as ``SMPLHGenerator.get_smplh`` hands it to the fitters, without its
    constructor (which reads the licensed model file): pose / betas / trans / offsets Parameters + the synthetic SMPL-H layer."""
    from lib_smpl.wrapper_pytorch import SMPLPyTorchWrapperBatch
    from lib_smpl.body_landmark import load_regressors
    S = SMPLPyTorchWrapperBatch.__new__(SMPLPyTorchWrapperBatch); torch.nn.Module.__init__(S)
    P = torch.nn.Parameter
    S.model_root, S.hands, S.device, S.gender = "synthetic", True, "cpu", "male"
    S.betas, S.pose, S.trans = P(d["betas"].clone()), P(d["pose"].clone()), P(d["trans"].clone())
    S.offsets = P(torch.zeros(B, 6890, 3)); S.smpl = L; S.faces = torch.zeros(1, 3, dtype=torch.long)
    S.body25_reg_torch, S.face_reg_torch, S.hand_reg_torch = load_regressors("assets", batch_size=B)
    return S


def _real_split(L):
    """Let the reference's own ``split_smpl`` -> ``SMPLPyTorchWrapperBatchSplitParams.from_smpl`` run (lib_smpl/wrapper_pytorch.py:206-226):
    only the two things its constructor fetches from licensed / configured paths are redirected (the SMPL layer, the regressor folder)."""
    import lib_smpl.wrapper_pytorch as W
    from lib_smpl.body_landmark import load_regressors
    W.SMPL_Layer = lambda **kw: L
    W.load_regressors = lambda root, batch_size=5: load_regressors("assets", batch_size=batch_size)


def recon_loop_goldens(out_dir: str):
    """The SMPL refinement LOOP of the reference -- ReconFitterBehave.optimize_smpl (recon/recon_fit_behave.py:393-465: split_smpl, phase
    schedule, the two Adam set-ups, the decay, the early-stop rule on fp32 tensors, get_smpl_height, copy_smpl_params) -- executed UNPATCHED on
    the CPU on the seeded problem, starting from the un-split container the driver passes.  Two runs: 'a' = 1 + 1 + 1 + 2 outer iterations of 3
    steps (stops at step 11), 'b' = 1 + 1 + 1 + 12 outer iterations of 2 steps.  Per step: every loss term and the total; at the end the
    parameters of the RETURNED container, the height ratio, and whether the split parameters aliased the caller's storage (they do: from_smpl
    wraps views of smpl.pose.data / betas.data, so the other-betas updates survive copy_smpl_params).  -> recon_loop.npz"""
    d, net, make_smpl, F, B = _recon_setup()
    L = make_smpl().smpl
    _real_split(L)
    qd = {"crop_center": d["crop"], "body_center": d["body_center"]}
    names = ["df_h", "pose", "hand", "part", "pinit", "j2d", "stemp"]
    out = {"term_names": np.array(names)}
    for tag, kw in (("a", dict(steps_per_iter=3, max_iter=2)), ("b", dict(steps_per_iter=2, max_iter=12))):
        S = _unsplit_container(d, L, B)
        dd = {"part_labels": d["labels"][None].repeat(B, 1), "net": net, "query_dict": qd, "pose_init": d["pose_init"], "body_kpts": d["body_kpts"]}
        hist, terms, splits = [], [], []
        orig_sum, orig_fwd, orig_split = F.sum_dict, F.forward_smpl, F.split_smpl

        def recording_sum(loss_dict, weight_dict, it):
            v = orig_sum(loss_dict, weight_dict, it)
            hist.append(float(v))
            return v

        def recording_fwd(smpl, data_dict, phase):
            ld = orig_fwd(smpl, data_dict, phase)
            terms.append([float(ld[k]) if k in ld else np.nan for k in names])
            return ld

        def recording_split(smpl):
            sp = orig_split(smpl)
            splits.append(sp)
            return sp
        F.sum_dict, F.forward_smpl, F.split_smpl = recording_sum, recording_fwd, recording_split
        try:
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                S2, scale = F.optimize_smpl(S, dd, iter_for_betas=1, iter_for_pose=1, iter_for_kpts=1, **kw)
        finally:
            del F.sum_dict, F.forward_smpl, F.split_smpl
        sp = splits[0]
        alias = bool(sp.other_betas.data_ptr() == S.betas.data[:, 2:].data_ptr() and sp.body_pose.data_ptr() == S.pose.data[:, 3:].data_ptr())
        h = np.array(hist, np.float64)
        # distance of every eligible early-stop test from its threshold (a knife-edge golden would be a useless golden)
        first_ok = (0.25 * kw["max_iter"] + 2)
        margin = []
        for i in range(1, len(h)):
            it = i // kw["steps_per_iter"]
            if it > first_ok:
                margin.append((abs(h[i - 1] - h[i]) / h[i - 1]) / (h[i - 1] * 1e-3))
        out.update({f"{tag}_hist": h, f"{tag}_terms": np.array(terms, np.float64), f"{tag}_pose": S2.pose.detach().numpy().copy(),
                    f"{tag}_betas": S2.betas.detach().numpy().copy(), f"{tag}_trans": S2.trans.detach().numpy().copy(),
                    f"{tag}_scale": scale.detach().numpy().copy(), f"{tag}_alias": np.array(alias), f"{tag}_stop_ratio": np.array(margin),
                    f"{tag}_betas_changed": np.array(float((S2.betas.detach() - d["betas"]).abs()[:, 2:].max()))})
        print(f"recon_loop[{tag}]: steps", len(hist), "of", (3 + kw["max_iter"]) * kw["steps_per_iter"], "losses", hist[0], "->", hist[-1], "alias", alias,
              "other-betas moved by", float(out[f"{tag}_betas_changed"]), "stop ratios (x threshold)", np.round(margin, 3))
    np.savez_compressed(os.path.join(out_dir, "recon_loop.npz"), **out)


def _install_sil_stubs():
    """Third-party operators of SilLossROI that cannot be installed, restated so that the reference's OWN class runs: detectron2's
    BitMasks.crop_and_resize (aligned RoIAlign of the boolean mask, >= 0.5) and BoxMode.convert on torchvision / numpy, neural_renderer's
    Renderer(mode='silhouettes') + its pseudo-gradient on oracle/raster_ref.py (parity unpinned, see its header).  cv2 and scipy are real."""
    from oracle import raster_ref as RR
    from torchvision.ops import roi_align

    class BitMasks:
        def __init__(self, t): self.tensor = t
        def crop_and_resize(self, boxes, size):
            n = self.tensor.shape[0]
            rois = torch.cat([torch.arange(n, dtype=torch.float32)[:, None], boxes.float()], 1)
            return roi_align((self.tensor != 0).float()[:, None], rois, (size, size), 1.0, 0, True)[:, 0] >= 0.5

    class BoxMode:
        XYXY_ABS, XYWH_ABS = 0, 1
        @staticmethod
        def convert(box, from_mode, to_mode):
            b = np.array(box, dtype=float).copy()
            if from_mode == to_mode:
                return b
            if from_mode == BoxMode.XYXY_ABS:
                b[:, 2:] -= b[:, :2]
            else:
                b[:, 2:] += b[:, :2]
            return b

    class _SilFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, verts, faces, K, size):
            v, f = verts.detach().double().numpy(), faces.numpy()
            imgs, saved = [], []
            for b in range(v.shape[0]):
                K4 = (float(K[b, 0, 0]), float(K[b, 1, 1]), float(K[b, 0, 2]), float(K[b, 1, 2]))
                fv = RR.faces_of(RR.project(v[b], K4), f)
                idx, alpha, _ = RR.rasterize_fast(fv, size)
                imgs.append(alpha); saved.append((fv, idx, alpha, K4))
            ctx.saved, ctx.v, ctx.f, ctx.size = saved, v, f, size
            return torch.from_numpy(np.stack(imgs)).float()

        @staticmethod
        def backward(ctx, g):
            g = g.double().numpy()
            out = []
            for b, (fv, idx, alpha, K4) in enumerate(ctx.saved):
                gf = RR.backward_faces(fv, idx, alpha, g[b], ctx.size)
                out.append(RR.backward_verts(gf, ctx.v[b], ctx.f, K4))
            return torch.from_numpy(np.stack(out)).float(), None, None, None

    class Renderer:
        def __init__(self, image_size=256, K=None, R=None, t=None, orig_size=1, anti_aliasing=False, **kw):
            assert orig_size == 1 and not anti_aliasing
            self.image_size, self.K = image_size, K
        def __call__(self, verts, faces, mode="silhouettes"):
            assert mode == "silhouettes"
            return _SilFn.apply(verts, faces[0], self.K, self.image_size)

    nr = _stub("neural_renderer", Renderer=Renderer)
    nr.renderer = _stub("neural_renderer.renderer", Renderer=Renderer)
    return BitMasks, BoxMode


def recon_obj_loop_goldens(out_dir: str):
    """The object / joint optimisation LOOP of the reference -- ReconFitterTriVisFull.optimize_smpl_object (recon/recon_fit_trivis_full.py:283-377)
    with its own SilLossROI construction (recon/obj_pose_roi.py:21-107), split_smpl, the three optimisers, phase switches ('object only' 2 outer
    iterations -> 'sil' 2 -> 'joint' up to 101, max_iter is hard-wired to 100), the per-phase decay, decopose_axis noise (torch.rand replaced by a
    seeded sequence that the GPU test replays), the contact sets computed once on the first joint step, and the joint-phase early stop -- run on the
    CPU, one step per outer iteration.  Third-party pieces restated: see _install_sil_stubs.  -> recon_obj_loop.npz"""
    d, net, make_smpl, F, B = _recon_setup()
    BitMasks, BoxMode = _install_sil_stubs()
    import recon.bbox as BB
    import recon.obj_pose_roi as OPR
    OPR.BitMasks, BB.BoxMode = BitMasks, BoxMode
    OPR.nr = sys.modules["neural_renderer"]
    import recon.recon_fit_trivis_full as M
    M.SilLossROI = OPR.SilLossROI
    OPR.SilLossROI.__init__.__defaults__ = tuple("cpu" if x == "cuda:0" else x for x in OPR.SilLossROI.__init__.__defaults__)
    torch.cuda.FloatTensor = torch.FloatTensor
    from recon_problem import make_loop_extras
    from argparse import Namespace
    e = make_loop_extras(d)
    L = make_smpl().smpl
    _real_split(L)
    F.scan = Namespace(v=e["temp_v"].astype(np.float64), f=e["temp_f"])
    F.get_opt_iters = lambda: {"sil": 2, "object": 2}
    qd = {"crop_center": d["crop"], "body_center": d["body_center"]}
    names = ["otemp", "ovtemp", "mask", "scale", "trans", "object", "ocent", "contact"]
    S = _unsplit_container(d, L, B)
    R_, t_ = d["obj_R"].clone().requires_grad_(True), d["obj_t"].clone().requires_grad_(True)
    dd = {"images": e["images_sil"], "query_dict": qd, "camera_params": {}, "crop_size": 1200, "net_input_size": d["images"].shape[-1],
          "smpl": S, "obj_R": R_, "obj_t": t_, "obj_s": d["obj_s"].clone(), "objects": d["objects"], "occ_ratios": d["occ"]}
    hist, terms, phases, draws = [], [], [], [0]
    orig_sum, orig_fwd = F.sum_dict, F.forward_step

    def recording_sum(loss_dict, weight_dict, it):
        v = orig_sum(loss_dict, weight_dict, it)
        hist.append(float(v))
        return v

    def recording_fwd(model, smpl, data_dict, obj_R, obj_t, obj_s, phase):
        ld = orig_fwd(model, smpl, data_dict, obj_R, obj_t, obj_s, phase)
        terms.append([float(ld[k]) if k in ld else np.nan for k in names])
        phases.append(phase)
        return ld
    F.sum_dict, F.forward_step = recording_sum, recording_fwd
    real_rand = torch.rand

    def seeded_rand(*a, **k):
        assert tuple(a) == (B, 3, 3)
        draws[0] += 1
        return e["noise_seq"][draws[0] - 1].clone()
    torch.rand = seeded_rand
    try:
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            S2, R2, t2 = F.optimize_smpl_object(net, dd, joint_iter=1, steps_per_iter=1)
    finally:
        torch.rand = real_rand
        del F.sum_dict, F.forward_step
    sil = dd["silhouette"]
    h = np.array(hist, np.float64)
    ratios = [(abs(h[i - 1] - h[i]) / h[i - 1]) / (h[i - 1] * 1e-4) for i in range(26, len(h))]
    out = {"term_names": np.array(names), "hist": h, "terms": np.array(terms, np.float64), "phases": np.array(phases), "n_draws": np.array(draws[0]),
           "obj_R": R2.detach().numpy().copy(), "obj_t": t2.detach().numpy().copy(), "rot_final": M.ReconFitterTriVisFull.decopose_axis(R2.detach(), no_rand=True).numpy(),
           "rot_init": dd["rot_init"].numpy().copy(), "trans_init": dd["trans_init"].numpy().copy(), "smpl_center": dd["smpl_center"].numpy().copy(),
           "df_obj_h": dd["df_obj_h"].numpy().copy(), "df_hum_o": dd["df_hum_o"].numpy().copy(), "parts_obj": dd["parts_obj"].numpy().copy(),
           "keep_mask": sil.keep_mask.numpy().copy(), "image_ref": sil.image_ref.numpy().copy(), "K_roi": sil.renderer.K.numpy().copy(),
           "stop_ratio": np.array(ratios)}
    np.savez_compressed(os.path.join(out_dir, "recon_obj_loop.npz"), **out)
    n_h, n_o = int((dd["df_hum_o"] < 0.08).sum()), int((dd["df_obj_h"] < 0.08).sum())
    print("recon_obj_loop.npz: steps", len(hist), "losses", hist[0], "->", hist[-1], "draws", draws[0], "contact verts (human, object)", n_h, n_o,
          "contact term present in", int(np.isfinite(out["terms"][:, names.index("contact")]).sum()), "steps; stop ratios (x threshold)", np.round(ratios[-5:], 3))


def driver_goldens(out_dir: str):
    """Host arithmetic of the fit_recon driver, from the reference's own methods called unbound on the CPU: ReconFitterBase.scale_body_kpts
    (recon/recon_fit_base.py:397-409) and ReconFitterBehave.combine_mini_batches (recon/recon_fit_behave.py:152-183).  -> driver_small.npz"""
    from argparse import Namespace
    for n in ("trimesh", "igl", "open3d", "zstd", "neural_renderer"):
        _stub(n)
    _stub("pytorch3d"); _stub("pytorch3d.loss", chamfer_distance=None)
    _stub("pytorch3d.structures", Pointclouds=object, Meshes=None); _stub("pytorch3d.ops", knn_points=None, sample_points_from_meshes=None)
    _stub("mesh_intersection"); _stub("mesh_intersection.bvh_search_tree", BVH=object); _stub("mesh_intersection.loss")
    _stub("detectron2"); _stub("detectron2.structures", BitMasks=object, BoxMode=object, Boxes=object); _stub("detectron2.structures.boxes", BoxMode=object)
    for m in list(sys.modules):
        if m.startswith(("psbody", "pytorch3d", "detectron2", "mesh_intersection", "skimage", "chumpy")):
            sys.modules[m].__path__ = []
    sys.modules["psbody.mesh"].MeshViewers = object
    import recon.recon_fit_base as MB                                                  # reference
    import recon.recon_fit_behave as MBH                                               # reference
    rng = np.random.default_rng(71)
    B = 4
    kpts = torch.from_numpy(np.concatenate([rng.uniform(0, 2048, (B, 25, 2)), rng.random((B, 25, 1))], -1).astype(np.float32))
    rs, cs = torch.from_numpy(rng.uniform(0.8, 1.3, B).astype(np.float32)), torch.from_numpy(rng.uniform(0.7, 1.4, B).astype(np.float32))
    cc = torch.from_numpy(rng.uniform(600, 1400, (B, 2)).astype(np.float32))
    shim = Namespace(camera=Namespace(crop_size=1200), net_in_size=512)
    out = {"kpts": kpts.numpy(), "resize_scale": rs.numpy(), "crop_scale": cs.numpy(), "crop_center": cc.numpy(),
           "scaled": MB.ReconFitterBase.scale_body_kpts(shim, kpts, rs, cs, cc).numpy(),
           "scaled_unit": MB.ReconFitterBase.scale_body_kpts(shim, kpts, torch.ones(B), torch.ones(B), cc).numpy()}
    mk = lambda b, n: {t: {"points": torch.from_numpy(rng.standard_normal((b, n, 3)).astype(np.float32)), "parts": torch.from_numpy(rng.integers(0, 14, (b, n))).float(),
                            "centers": torch.from_numpy(rng.standard_normal((b, 6)).astype(np.float32)), "pca_axis": torch.from_numpy(rng.standard_normal((b, 3, 3)).astype(np.float32)),
                            "visibility": torch.from_numpy(rng.random((b,)).astype(np.float32))} for t in ("human", "object")}
    pcs = [mk(2, 50), mk(1, 40), mk(2, 47)]
    comb = MBH.ReconFitterBehave.combine_mini_batches(None, pcs, 40)
    for i, pc in enumerate(pcs):
        for t in pc:
            for k, v in pc[t].items():
                out[f"pc{i}.{t}.{k}"] = v.numpy()
    for t in comb:
        out[f"comb.{t}.order"] = np.array(list(comb[t]))
        for k, v in comb[t].items():
            out[f"comb.{t}.{k}"] = v.numpy()
    np.savez_compressed(os.path.join(out_dir, "driver_small.npz"), **out)
    print("driver_small.npz:", out["scaled"].shape, {k: out[k].shape for k in out if k.startswith("comb.human.") and not k.endswith("order")})


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    import_reference(a.ref)
    torch.set_num_threads(os.cpu_count())
    if a.only in ("", "sifnet"):
        sifnet_goldens(HERE)
    if a.only in ("", "smpl"):
        smpl_goldens(HERE)
    if a.only in ("", "assets"):
        asset_fixtures(HERE, a.ref)
    if a.only in ("", "fit"):
        fit_smplt_goldens(HERE)
    if a.only in ("", "recon"):
        recon_goldens(HERE)
    if a.only in ("", "eval"):
        eval_goldens(HERE)
    if a.only == "smooth":                  # stubs `behave` / `yacs`: run on its own
        smooth_goldens(HERE)
    if a.only == "frameio":                 # stubs the `data` package (real cv2): run on its own
        frameio_goldens(HERE)
    if a.only == "interp":                  # stubs `behave`: run on its own
        interp_goldens(HERE)
    if a.only == "generator":               # stubs `trainer`: run on its own
        generator_goldens(HERE)
    if a.only == "render":                  # stubs `cv2` / `behave` / `neural_renderer`: run on its own
        render_goldens(HERE)
    if a.only == "roi":                     # stubs `cv2` / `detectron2` / `neural_renderer`: run on its own
        roi_goldens(HERE)
    if a.only == "evalseq":                 # replaces psbody's Mesh and several loaders: run on its own
        evalseq_goldens(HERE)
    if a.only == "io":                      # replaces psbody's Mesh: run on its own
        io_goldens(HERE)
    if a.only == "pack":                    # stubs `behave.frame_data` / `lib_smpl`: run on its own
        pack_goldens(HERE)
    if a.only == "reconloop":
        recon_loop_goldens(HERE)
    if a.only == "reconobjloop":            # replaces SilLossROI's third-party operators: run on its own
        recon_obj_loop_goldens(HERE)
    if a.only == "driver":
        driver_goldens(HERE)
    if a.only == "infill_io":
        infill_io_goldens(HERE)
    if a.only == "infill":                  # stubs `behave` / `trainer`: run on its own
        infill_goldens(HERE)
