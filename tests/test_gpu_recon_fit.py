"""Loss assembly of the joint optimisation on the B200 against goldens produced by the UNMODIFIED reference methods
``forward_smpl`` / ``forward_step`` (tests/golden/recon_small.npz, make_golden.py --only recon)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from recon_problem import B, make_problem
from vistracker_b200 import CHORETriplaneVisibility, default_options, resolve_dims
from vistracker_b200.synth import synthetic_state_dict

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def ctx():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from vistracker_b200.recon_fit import Priors, ReconFitterTriVisFull, SMPLParams
    from vistracker_b200.smpl import LandmarkRegressor, SMPL_Layer
    d = make_problem()
    dims = resolve_dims(default_options())
    net = CHORETriplaneVisibility(default_options(), device="cuda:0").eval()
    net.load_state_dict(synthetic_state_dict(dims, seed=0))
    net.filter(d["images"].cuda())
    layer = SMPL_Layer.from_buffers(d["model"], d["model"]["parents"], "cuda:0")
    reg = LandmarkRegressor(np.stack([d["reg"][0], d["reg"][1]]), d["reg"][2], d["reg"][3], "cuda:0")
    fitter = ReconFitterTriVisFull(net, Priors(d["assets"], "cuda:0"), d["labels"])
    make_smpl = lambda: SMPLParams(layer, reg, d["pose"], d["betas"], d["trans"])
    return d, fitter, make_smpl


def _close(a, b, tol=TOL):
    return abs(a - b) <= tol * max(abs(b), 1e-6)


def test_forward_smpl_matches_reference(ctx, golden):
    d, fitter, make_smpl = ctx
    g = golden("recon_small.npz")
    smpl = make_smpl()
    c = lambda t: t.cuda()
    dd = {"part_labels": c(d["labels"])[None].repeat(B, 1), "query_dict": {"crop_center": c(d["crop"]), "body_center": c(d["body_center"])},
          "pose_init": c(d["pose_init"]), "body_kpts": c(d["body_kpts"])}
    ld = fitter.forward_smpl(smpl, dd, "kpts")
    assert list(ld) == ["df_h", "pose", "hand", "part", "pinit", "j2d", "stemp"]
    for k, v in ld.items():
        assert _close(float(v), float(g[f"smpl_{k}"])), (k, float(v), float(g[f"smpl_{k}"]))
    fitter.sum_dict(ld, fitter.get_loss_weights(), 2 / 3).backward()
    assert rel_err(torch.cat([smpl.global_pose.grad, smpl.body_pose.grad], 1).cpu(), g["smpl_g_pose"]) < TOL
    assert rel_err(torch.cat([smpl.top_betas.grad, smpl.other_betas.grad], 1).cpu(), g["smpl_g_betas"]) < TOL
    assert rel_err(smpl.trans.grad.cpu(), g["smpl_g_trans"]) < TOL


@pytest.mark.parametrize("phase,tag,decay", [("object only", "obj", 1), ("joint", "joint", 4 / 3)])
def test_forward_step_matches_reference(ctx, golden, phase, tag, decay):
    d, fitter, make_smpl = ctx
    g = golden("recon_small.npz")
    c = lambda t: t.cuda()
    R_, t_ = c(d["obj_R"]).requires_grad_(True), c(d["obj_t"]).requires_grad_(True)
    dd = {"objects": c(d["objects"]), "query_dict": {"crop_center": c(d["crop"]), "body_center": c(d["body_center"])},
          "occ_ratios": c(d["occ"]), "smpl_center": c(d["smpl_center"]), "df_obj_h": c(d["df_obj_h"]), "df_hum_o": c(d["df_hum_o"]),
          "parts_obj": c(d["parts_obj"])}
    ld = fitter.forward_step(make_smpl(), dd, R_, t_, c(d["obj_s"]), phase, noise=c(d["noise"]))
    expect = [k[len(tag) + 1:] for k in g if k.startswith(tag + "_") and not k.startswith(tag + "_g_")]
    assert list(ld) == expect
    for k, v in ld.items():
        assert _close(float(v), float(g[f"{tag}_{k}"])), (k, float(v), float(g[f"{tag}_{k}"]))
    fitter.sum_dict(ld, fitter.get_loss_weights(), decay).backward()
    # the golden's d/dR goes through torch.svd's fp32 backward on the CPU; ours is a closed form evaluated in fp64
    assert rel_err(R_.grad.cpu(), g[f"{tag}_g_R"]) < 3e-4
    assert rel_err(t_.grad.cpu(), g[f"{tag}_g_t"]) < TOL


def test_optimisation_loops_run_and_reduce_the_loss(ctx):
    """Short runs of both loops (phase switches included through tiny iteration counts)."""
    from vistracker_b200.render import SilLossROI
    d, fitter, make_smpl = ctx
    c = lambda t: t.cuda()
    qd = {"crop_center": c(d["crop"]), "body_center": c(d["body_center"])}
    smpl = make_smpl()
    dd = {"part_labels": c(d["labels"])[None].repeat(B, 1), "query_dict": qd, "pose_init": c(d["pose_init"]), "body_kpts": c(d["body_kpts"])}
    smpl, hist = fitter.optimize_smpl(smpl, dd, 1, 1, 1, steps_per_iter=3, max_iter=2)
    # the reference's early-stop rule |prev - loss| / prev < prev * 1e-3 scales with the loss value, so it may fire once
    # it > 0.25 * max_iter + 2; the run is 15 steps at most
    assert 10 <= len(hist) <= 15 and np.isfinite(hist).all() and min(hist[5:]) < hist[1]
    # object: template = the ellipsoid points' convex hull, ROI = whole crop
    from scipy.spatial import ConvexHull
    tmpl = d["objects"][0].numpy()
    faces = ConvexHull(tmpl).simplices
    K = SilLossROI.compute_K_roi((424.0, 168.0, 1200.0, 1200.0))[None].repeat(B, 1, 1)
    ref = torch.zeros(B, 64, 64); ref[:, 20:44, 24:40] = 1
    sil = SilLossROI(torch.ones(B, 64, 64), ref, K, tmpl, faces, rend_size=64)
    R_, t_ = c(d["obj_R"]).requires_grad_(True), c(d["obj_t"]).requires_grad_(True)
    dd = {"objects": c(d["objects"]), "query_dict": qd, "occ_ratios": c(d["occ"]), "obj_R": R_, "obj_t": t_, "obj_s": c(d["obj_s"]),
          "silhouette": sil}
    fitter.get_opt_iters = staticmethod(lambda: {"sil": 2, "object": 2})
    _, R_out, t_out, hist = fitter.optimize_smpl_object(smpl, dd, joint_iter=1, steps_per_iter=2, max_iter=1)
    assert 8 <= len(hist) <= 12 and np.isfinite(hist).all()          # 6 outer x 2 steps unless the joint-phase early stop fires
    assert "trans_init" in dd and "df_obj_h" in dd
    Rf = fitter.final_rotation(R_out)
    assert rel_err((Rf @ Rf.transpose(1, 2)).cpu(), torch.eye(3).expand(B, 3, 3)) < 1e-5


@pytest.mark.xfail(strict=False, reason="added after the round's GPU budget was spent: first B200 run of this comparison happens at round end")
def test_optimize_smpl_loop_follows_the_reference_loop(ctx, golden):
    """The whole SMPL refinement loop (phase schedule, both Adam set-ups, decay, early stop) against the reference's own
    ReconFitterBehave.optimize_smpl executed on the CPU (tests/golden/recon_loop.npz: 1 + 1 + 1 + 2 outer iterations of 3 steps)."""
    d, fitter, make_smpl = ctx
    g = golden("recon_loop.npz")
    c = lambda t: t.cuda()
    dd = {"part_labels": c(d["labels"])[None].repeat(B, 1), "query_dict": {"crop_center": c(d["crop"]), "body_center": c(d["body_center"])},
          "pose_init": c(d["pose_init"]), "body_kpts": c(d["body_kpts"])}
    smpl, hist = fitter.optimize_smpl(make_smpl(), dd, iter_for_betas=1, iter_for_pose=1, iter_for_kpts=1, steps_per_iter=3, max_iter=2)
    assert len(hist) == len(g["hist"])                                            # same early stop
    assert rel_err(np.asarray(hist), g["hist"]) < 1e-3
    assert rel_err(smpl.pose.detach().cpu(), g["pose"]) < 1e-3 and rel_err(smpl.trans.detach().cpu(), g["trans"]) < 1e-3
    assert rel_err(smpl.betas.detach().cpu(), g["betas"]) < 1e-3
