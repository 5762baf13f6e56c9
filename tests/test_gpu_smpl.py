"""SMPL-H layer and landmark regressors on the B200 against the reference golden vectors and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_err
from oracle.smpl_ref import landmarks, smpl_forward
from vistracker_b200.synth_smpl import synthetic_motion, synthetic_smplh, synthetic_smplh_surface

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def layer():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from vistracker_b200.smpl import SMPL_Layer
    model = synthetic_smplh(seed=3)
    return SMPL_Layer.from_buffers(model, model["parents"], "cuda:0"), model


def test_forward_backward_match_reference_golden(layer, golden):
    L, _ = layer
    g = golden("smpl_small.npz")
    pose, betas, trans = synthetic_motion(5, seed=5)
    pose[0, 3:6] = 0.0
    pose, betas, trans = (t.cuda().requires_grad_(True) for t in (pose, betas, trans))
    verts, jtr, v_posed, naked = L(pose, th_betas=betas, th_trans=trans, th_offsets=torch.zeros(5, 6890, 3, device="cuda"))
    assert rel_err(verts.detach().cpu(), g["verts"]) < TOL
    assert rel_err(jtr.detach().cpu(), g["jtr"]) < TOL
    assert rel_err(v_posed.detach().cpu(), g["v_posed"]) < TOL
    rng = np.random.Generator(np.random.PCG64(17))
    gv = torch.from_numpy(rng.standard_normal(tuple(verts.shape), dtype=np.float32)).cuda()
    gj = torch.from_numpy(rng.standard_normal(tuple(jtr.shape), dtype=np.float32)).cuda()
    ((verts * gv).sum() + (jtr * gj).sum()).backward()
    assert rel_err(pose.grad.cpu(), g["g_pose"]) < TOL
    assert rel_err(betas.grad.cpu(), g["g_betas"]) < TOL
    assert rel_err(trans.grad.cpu(), g["g_trans"]) < TOL


@pytest.mark.parametrize("B,scale,with_offsets", [(1, 1.0, False), (37, 1.0, False), (96, 0.9, True)])
def test_matches_oracle_fp64(layer, B, scale, with_offsets):
    L, model = layer
    pose, betas, trans = synthetic_motion(B, seed=100 + B)
    off = torch.randn(B, 6890, 3, generator=torch.Generator().manual_seed(B)) * 0.01 if with_offsets else None
    ref_in = [t.double().requires_grad_(True) for t in (pose, betas, trans)]
    rv, rj, rvp, rn = smpl_forward(model, *ref_in, None if off is None else off.double(), scale)
    ins = [t.cuda().requires_grad_(True) for t in (pose, betas, trans)]
    verts, jtr, v_posed, naked = L(*ins, th_offsets=None if off is None else off.cuda(), scale=scale)
    for a, b in ((verts, rv), (jtr, rj), (v_posed, rvp), (naked, rn)):
        assert rel_err(a.detach().cpu(), b.detach()) < 1e-5
    gen = torch.Generator().manual_seed(7)
    gv, gj = torch.randn(B, 6890, 3, generator=gen), torch.randn(B, 52, 3, generator=gen)
    ((rv * gv.double()).sum() + (rj * gj.double()).sum()).backward()
    ((verts * gv.cuda()).sum() + (jtr * gj.cuda()).sum()).backward()
    for a, b in zip(ins, ref_in):
        assert rel_err(a.grad.cpu(), b.grad) < TOL
    # only-joints and only-vertices cotangents (the fitters use both forms)
    ins2 = [t.cuda().requires_grad_(True) for t in (pose, betas, trans)]
    L(*ins2, scale=scale)[1].sum().backward()
    ref2 = [t.double().requires_grad_(True) for t in (pose, betas, trans)]
    smpl_forward(model, *ref2, None, scale)[1].sum().backward()
    for a, b in zip(ins2, ref2):
        assert rel_err(a.grad.cpu(), b.grad) < TOL


def test_surface_skinned_body_matches_oracle_fp64():
    """The human-shaped model (skinning follows the surface: the lanes of a warp share their joints, so smpl_skin_bwd_kernel sums whole groups with
    a butterfly before ONE lane touches the shared accumulators -- the random model above only takes the direct-atomics path) against the fp64
    restatement, forward and backward, B = 40 with a gradient on the vertices only and on vertices + joints."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from vistracker_b200.smpl import SMPL_Layer
    model = synthetic_smplh_surface(seed=3)
    L = SMPL_Layer.from_buffers(model, model["parents"], "cuda:0")
    B = 40
    pose, betas, trans = synthetic_motion(B, seed=321)
    for with_joints in (False, True):
        ref_in = [t.double().requires_grad_(True) for t in (pose, betas, trans)]
        rv, rj, _, _ = smpl_forward(model, *ref_in, None, 1.0)
        ins = [t.cuda().requires_grad_(True) for t in (pose, betas, trans)]
        verts, jtr, _, _ = L(*ins)
        assert rel_err(verts.detach().cpu(), rv.detach()) < 1e-5 and rel_err(jtr.detach().cpu(), rj.detach()) < 1e-5
        gen = torch.Generator().manual_seed(11)
        gv, gj = torch.randn(B, 6890, 3, generator=gen), torch.randn(B, 52, 3, generator=gen) * float(with_joints)
        ((rv * gv.double()).sum() + (rj * gj.double()).sum()).backward()
        ((verts * gv.cuda()).sum() + (jtr * gj.cuda()).sum()).backward()
        for a, b in zip(ins, ref_in):
            assert rel_err(a.grad.cpu(), b.grad) < TOL


def test_body25_landmarks_with_the_reference_asset(layer):
    """assets/body25_regressor.pkl is the one numeric fixture of this path the reference ships;
    a copy of its COO triplet is committed under tests/golden/ (make_golden.py)."""
    from vistracker_b200.smpl import LandmarkRegressor
    L, model = layer
    a = np.load(os.path.join(ROOT, "tests", "golden", "assets.npz"))
    idx, val, shape = np.stack([a["body25_row"], a["body25_col"]]), a["body25_val"], a["body25_shape"]
    assert tuple(shape) == (6890, 25) and val.shape == (8481,)
    reg = LandmarkRegressor(idx, val, shape, "cuda:0")
    verts = torch.randn(7, 6890, 3, generator=torch.Generator().manual_seed(1))
    ref_in = verts.double().requires_grad_(True)
    ref = landmarks(torch.as_tensor(idx).long(), torch.as_tensor(val), tuple(shape), ref_in)
    x = verts.cuda().requires_grad_(True)
    out = reg(x)
    assert out.shape == (7, 25, 3) and rel_err(out.detach().cpu(), ref.detach()) < 1e-5
    g = torch.randn(7, 25, 3, generator=torch.Generator().manual_seed(2))
    (ref * g.double()).sum().backward(); (out * g.cuda()).sum().backward()
    assert rel_err(x.grad.cpu(), ref_in.grad) < 1e-5
