"""The SMPL-T fit oracle against the golden produced with the reference's own compute_loss / sum_dict / Adam set-up."""
import numpy as np
import torch

from conftest import rel_err
from fit_problem import load_assets, synthetic_fit_problem
from oracle import fit_smplt_ref as F


def test_loss_terms_and_gradients_match_reference(golden):
    g = golden("fit_smplt_small.npz")
    a, reg = load_assets()
    model, kpts, pose0, betas0, trans0 = synthetic_fit_problem(12, seed=9)
    pose, betas, trans = (t.clone().requires_grad_(True) for t in (pose0, betas0, trans0))
    ld = F.compute_loss(model, reg, a, pose, betas, trans, kpts, pose0.clone())
    for k, v in ld.items():
        assert abs(float(v) - float(g[f"loss0_{k}"])) <= 2e-5 * abs(float(g[f"loss0_{k}"])), k
    F.total_loss(ld, 0).backward()
    assert rel_err(pose.grad[:, :66], g["g0_pose"]) < 1e-4
    assert rel_err(betas.grad, g["g0_betas"]) < 1e-4
    assert rel_err(trans.grad, g["g0_trans"]) < 1e-4


def test_hundred_adam_steps_follow_the_reference_trajectory(golden):
    g = golden("fit_smplt_small.npz")
    a, reg = load_assets()
    model, kpts, pose0, betas0, trans0 = synthetic_fit_problem(12, seed=9)
    pose, betas, trans, losses, snaps = F.fit(model, reg, a, pose0, betas0, trans0, kpts, n_outer=10, record=(1, 10, 80, 81, 100))
    assert rel_err(np.array(losses), g["losses"]) < 1e-4
    for s in (1, 10, 80, 81, 100):
        assert rel_err(snaps[s][0], g[f"pose_{s}"]) < 1e-4, s
        assert rel_err(snaps[s][1], g[f"betas_{s}"]) < 1e-4, s
        assert rel_err(snaps[s][2], g[f"trans_{s}"]) < 1e-4, s
    assert torch.equal(snaps[100][0][:, 66:], pose0[:, 66:])          # hand pose is never optimised
