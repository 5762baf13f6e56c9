"""SMPL-T pre-fit step on the B200 against the golden produced with the reference's own compute_loss / sum_dict / Adam
(tests/golden/fit_smplt_small.npz) and against the CPU oracle on a second problem."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from fit_problem import load_assets, synthetic_fit_problem
from oracle import fit_smplt_ref as F

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def fitter():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from vistracker_b200.fit_smplt import SMPLHFitter30fps
    from vistracker_b200.smpl import LandmarkRegressor, SMPL_Layer
    a, reg = load_assets()
    model, *_ = synthetic_fit_problem(3, seed=1)
    layer = SMPL_Layer.from_buffers(model, model["parents"], "cuda:0")
    body25 = LandmarkRegressor(np.stack([reg[0], reg[1]]), reg[2], reg[3], "cuda:0")
    return SMPLHFitter30fps(layer, body25, a)


def test_loss_terms_and_gradients_match_reference(fitter, golden):
    g = golden("fit_smplt_small.npz")
    _, kpts, pose0, betas0, trans0 = synthetic_fit_problem(12, seed=9)
    losses, g_pose, g_betas, g_trans = fitter.compute_loss(pose0, betas0, trans0, kpts, decay=0)
    for k in ("kpts", "temp", "ptemp", "pose", "hand", "pinit"):
        assert abs(losses[k] - float(g[f"loss0_{k}"])) <= TOL * abs(float(g[f"loss0_{k}"])), k
    assert rel_err(g_pose[:, :66].cpu(), g["g0_pose"]) < TOL
    assert rel_err(g_betas.cpu(), g["g0_betas"]) < TOL
    assert rel_err(g_trans.cpu(), g["g0_trans"]) < TOL


@pytest.mark.parametrize("use_graph", [False, True])
def test_hundred_steps_follow_the_reference_trajectory(fitter, golden, use_graph):
    g = golden("fit_smplt_small.npz")
    _, kpts, pose0, betas0, trans0 = synthetic_fit_problem(12, seed=9)
    out = fitter.fit_batch(pose0, betas0, trans0, kpts, max_iter=10, early_stop=False, use_graph=use_graph, record=(1, 10, 80, 81, 100))
    assert out["steps"] == 100
    assert rel_err(out["losses"], g["losses"]) < TOL
    for s in (1, 10, 80, 81, 100):
        p, b, t = out["snapshots"][s]
        assert rel_err(p.cpu(), g[f"pose_{s}"]) < TOL, s
        assert rel_err(b.cpu(), g[f"betas_{s}"]) < TOL, s
        assert rel_err(t.cpu(), g[f"trans_{s}"]) < TOL, s
    assert torch.equal(out["pose"][:, 66:].cpu(), pose0[:, 66:])


def test_other_batch_size_matches_oracle_and_early_stop_is_honoured(fitter):
    a, reg = load_assets()
    model, kpts, pose0, betas0, trans0 = synthetic_fit_problem(40, seed=23)
    ref = F.fit(model, reg, a, pose0, betas0, trans0, kpts, n_outer=3)
    out = fitter.fit_batch(pose0, betas0, trans0, kpts, max_iter=3, early_stop=False)
    assert rel_err(out["losses"], np.array(ref[3])) < TOL
    assert rel_err(out["pose"].cpu(), ref[0]) < TOL and rel_err(out["trans"].cpu(), ref[2]) < TOL
    with pytest.raises(ValueError):
        fitter.fit_batch(pose0[:2], betas0[:2], trans0[:2], kpts[:2])


def test_smoothed_refit_schedule_matches_oracle(fitter):
    """SMPLHFitterSmoothed (preprocess/fit_SMPLH_smoothed.py): no global-pose phase, 30 outer iterations at most, all-pose Adam from step 1."""
    from vistracker_b200.fit_smplt import SMPLHFitterSmoothed
    a, reg = load_assets()
    model, kpts, pose0, betas0, trans0 = synthetic_fit_problem(24, seed=31)
    sm = SMPLHFitterSmoothed(fitter.smpl, fitter.reg, a)
    assert sm.get_globalopt_iters() == 0 and sm.get_max_iters() == 30
    ref = F.fit(model, reg, a, pose0, betas0, trans0, kpts, n_outer=4, iter_for_global=0)
    out = sm.fit_batch(pose0, betas0, trans0, kpts, max_iter=4, early_stop=False)
    assert out["steps"] == 40 and rel_err(out["losses"], np.array(ref[3])) < TOL
    assert rel_err(out["pose"].cpu(), ref[0]) < TOL and rel_err(out["betas"].cpu(), ref[1]) < TOL and rel_err(out["trans"].cpu(), ref[2]) < TOL
    assert float((out["pose"][:, 3:66].cpu() - pose0[:, 3:66]).abs().max()) > 0          # the body pose moves from the first step
