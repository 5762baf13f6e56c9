"""The SMPL-H oracle against golden vectors from the unmodified reference SMPL_Layer.forward (make_golden.py)."""
import numpy as np
import torch

from conftest import rel_err
from oracle.smpl_ref import rodrigues, smpl_forward
from vistracker_b200.synth_smpl import SMPLH_PARENTS, synthetic_motion, synthetic_smplh


def _inputs():
    pose, betas, trans = synthetic_motion(5, seed=5)
    pose[0, 3:6] = 0.0
    return pose, betas, trans


def test_smpl_forward_and_gradients_match_reference(golden):
    g = golden("smpl_small.npz")
    model = synthetic_smplh(seed=3)
    pose, betas, trans = _inputs()
    pose.requires_grad_(True); betas.requires_grad_(True); trans.requires_grad_(True)
    verts, jtr, v_posed, naked = smpl_forward(model, pose, betas, trans, torch.zeros(5, 6890, 3))
    assert rel_err(verts.detach(), g["verts"]) < 1e-5
    assert rel_err(jtr.detach(), g["jtr"]) < 1e-5
    assert rel_err(v_posed.detach(), g["v_posed"]) < 1e-5
    rng = np.random.Generator(np.random.PCG64(17))
    gv = torch.from_numpy(rng.standard_normal(tuple(verts.shape), dtype=np.float32))
    gj = torch.from_numpy(rng.standard_normal(tuple(jtr.shape), dtype=np.float32))
    ((verts * gv).sum() + (jtr * gj).sum()).backward()
    assert rel_err(pose.grad, g["g_pose"]) < 1e-4
    assert rel_err(betas.grad, g["g_betas"]) < 1e-4
    assert rel_err(trans.grad, g["g_trans"]) < 1e-4


def test_rodrigues_is_a_rotation_and_handles_zero():
    R = rodrigues(torch.tensor([[0.0, 0.0, 0.0], [0.3, -1.2, 0.7], [3.1, 0.0, 0.0]], dtype=torch.float64))
    eye = torch.eye(3, dtype=torch.float64)
    assert torch.allclose(R[0], eye, atol=1e-7)
    for r in R:
        assert torch.allclose(r @ r.t(), eye, atol=1e-12) and abs(float(torch.det(r)) - 1) < 1e-12


def test_kinematic_tree_is_topologically_ordered():
    assert SMPLH_PARENTS[0] == -1 and all(0 <= p < i for i, p in enumerate(SMPLH_PARENTS) if i)
