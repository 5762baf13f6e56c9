"""SLERP / LERP baseline (vistracker_b200/interp.py) against the reference's BaseInterpolator static methods (tests/golden/interp_small.npz)."""
import os

import numpy as np
import torch

from vistracker_b200 import interp as I

GOLD = os.path.join(os.path.dirname(__file__), "golden", "interp_small.npz")


def test_spans_slerp_and_lerp_match_reference():
    from scipy.spatial.transform import Rotation
    g = np.load(GOLD)
    occ = g["occ"]
    end_inds, start_inds = I.compute_missing_inds((occ < 0.3).astype(float))
    assert np.array_equal(end_inds, g["end_inds"]) and np.array_equal(start_inds, g["start_inds"]) and len(start_inds) > 5
    rot_q = Rotation.from_matrix(g["obj_angles_in"].transpose(0, 2, 1)).as_quat()
    q = I.interp_slerp(end_inds, rot_q, start_inds)
    assert q.shape == g["quat_out"].shape and np.abs(q - g["quat_out"]).max() < 1e-12
    tr = I.interp_lerp(end_inds, g["trans_in"].astype(np.float64), start_inds)
    assert np.abs(tr - g["trans_out"]).max() < 1e-12
    ang, tr2 = I.interpolate_object_rotations(torch.from_numpy(g["obj_angles_in"]), occ, 0.3, obj_trans=g["trans_in"].astype(np.float64))
    assert np.abs(ang - g["obj_angles_out"]).max() < 1e-12 and np.abs(tr2 - g["trans_out"]).max() < 1e-12


def test_no_occlusion_and_leading_span():
    rng = np.random.default_rng(0)
    from scipy.spatial.transform import Rotation
    R = Rotation.from_rotvec(rng.standard_normal((30, 3))).as_matrix().transpose(0, 2, 1)
    out = I.interpolate_object_rotations(R, np.ones(30), 0.3)
    assert np.abs(out - R).max() < 1e-12
    occ = np.ones(30); occ[:4] = 0.0; occ[10:15] = 0.0                              # a leading occluded span is left alone (reference: warning only)
    out = I.interpolate_object_rotations(R, occ, 0.3)
    assert out.shape == R.shape and np.abs(out[:9] - R[:9]).max() < 1e-12 and np.abs(out[15:] - R[15:]).max() < 1e-12
    assert np.abs(out[10:15] - R[10:15]).max() > 1e-3
    q0 = torch.tensor([[[1.0, 0, 0, 0]]], dtype=torch.float64)
    assert torch.equal(I.slerp(q0, q0.clone(), torch.tensor(0.5, dtype=torch.float64).reshape(1, 1, 1, 1)), q0.unsqueeze(1))
