"""oracle/recon_fit_ref.py (the CPU restatement used as checker and CPU baseline of the joint optimisation) pinned to the reference's OWN
loops: ReconFitterBehave.optimize_smpl and ReconFitterTriVisFull.optimize_smpl_object executed on the CPU by tests/golden/make_golden.py
(recon_loop.npz, recon_obj_loop.npz)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import recon_fit_ref as RF
from oracle import sifnet_ref as SR
from recon_problem import B, make_loop_extras, make_problem
from vistracker_b200 import default_options, resolve_dims
from vistracker_b200.synth import synthetic_state_dict

TOL = 1e-4


@pytest.fixture(scope="module")
def problem():
    torch.set_num_threads(8)
    d = make_problem()
    sd = synthetic_state_dict(resolve_dims(default_options()), seed=0)
    with torch.no_grad():
        maps = SR.sif_filter(sd, d["images"])
    P = RF.Problem(sd, maps, d["model"], d["reg"], d["assets"], d["labels"], d["crop"], d["body_center"])
    return d, P


def _close_terms(ours, ref):
    assert ours.shape == ref.shape and np.array_equal(np.isnan(ours), np.isnan(ref))
    m = ~np.isnan(ref)
    return float((np.abs(ours[m] - ref[m]) / np.maximum(np.abs(ref[m]), 1e-6)).max())


@pytest.mark.parametrize("tag,kw", [("a", dict(steps_per_iter=3, max_iter=2)), ("b", dict(steps_per_iter=2, max_iter=12))])
def test_optimize_smpl_restatement_equals_the_reference_loop(problem, golden, tag, kw):
    d, P = problem
    g = golden("recon_loop.npz")
    out = RF.optimize_smpl(P, d["pose"], d["betas"], d["trans"], d["pose_init"], d["body_kpts"], 1, 1, 1, **kw)
    assert len(out["hist"]) == len(g[f"{tag}_hist"]) and out["stopped"]
    assert rel_err(out["hist"], g[f"{tag}_hist"]) < TOL and _close_terms(out["terms"], g[f"{tag}_terms"]) < TOL
    for k in ("pose", "betas", "trans", "scale"):
        assert rel_err(out[k], g[f"{tag}_{k}"]) < TOL, k


def test_optimize_smpl_object_restatement_equals_the_reference_loop(problem, golden):
    d, P = problem
    g = golden("recon_obj_loop.npz")
    e = make_loop_extras(d)
    sil = RF.SilLoss(torch.from_numpy(g["keep_mask"]), torch.from_numpy(g["image_ref"]), torch.from_numpy(g["K_roi"]), e["temp_v"], e["temp_f"])
    draws = [0]

    def noise_fn():
        draws[0] += 1
        return e["noise_seq"][draws[0] - 1]
    out = RF.optimize_smpl_object(P, d["pose"], d["betas"], d["trans"], d["obj_R"], d["obj_t"], d["obj_s"], d["objects"], d["occ"], sil, noise_fn,
                                  it_obj=2, it_sil=2, joint_iter=1, steps_per_iter=1)
    assert len(out["hist"]) == len(g["hist"]) and draws[0] == int(g["n_draws"]) and list(out["phases"]) == list(g["phases"])
    assert list(g["term_names"]) == list(RF.OBJ_TERMS)
    assert rel_err(out["hist"], g["hist"]) < TOL and _close_terms(out["terms"], g["terms"]) < TOL
    assert rel_err(out["obj_R"], g["obj_R"]) < TOL and rel_err(out["obj_t"], g["obj_t"]) < TOL and rel_err(out["rot_final"], g["rot_final"]) < TOL
    st = out["state"]
    assert rel_err(st["df_obj_h"], g["df_obj_h"]) < TOL and rel_err(st["df_hum_o"], g["df_hum_o"]) < TOL
    assert rel_err(st["trans_init"], g["trans_init"]) < TOL and rel_err(st["rot_init"], g["rot_init"]) < TOL
