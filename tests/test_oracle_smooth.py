"""oracle/smooth_ref.py against the goldens produced by the reference's own SmoothNet modules and smoothers (smooth_small.npz)."""
import os

import numpy as np
import torch

from conftest import rel_err
from oracle import smooth_ref as S

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "smooth_small.npz"))


def _sd(prefix):
    return {k[len(prefix):]: torch.from_numpy(G[k]) for k in G.files if k.startswith(prefix)}


def test_rotation_conversions_match_reference():
    ax = G["conv_axis"]
    assert rel_err(S.axis_to_rot6d_np(ax), G["conv_np_rot6d"]) < 1e-6
    back = S.rot6d_to_axis(torch.from_numpy(G["conv_rot6d"]))
    assert rel_err(back, G["conv_axis_back"]) < 1e-5


def test_network_and_window_mean_match_reference():
    sd = _sd("smplt.")
    x = torch.from_numpy(G["input_data"]).permute(0, 2, 1)
    out = torch.cat([S.smoothnet_forward(sd, "pose_net.", x[:, :144]), x[:, 144:154], S.smoothnet_forward(sd, "trans_net.", x[:, 154:])], 1)
    assert rel_err(out.permute(0, 2, 1), G["denoised_clips"]) < 1e-5
    clips = S.seq2batches(torch.arange(90 * 2, dtype=torch.float32).reshape(90, 2), 64)
    assert torch.equal(S.slide_window_mean(clips), torch.arange(90 * 2, dtype=torch.float32).reshape(90, 2))


def test_smplt_smoother_matches_reference():
    poses, betas, trans = S.smooth_smplt(_sd("smplt."), G["poses_in"], G["betas_in"], G["trans_in"], int(G["W"]))
    assert rel_err(poses, G["poses_out"]) < 1e-4
    assert rel_err(betas, G["betas_out"]) < 1e-6 and rel_err(trans, G["trans_out"]) < 1e-5


def test_objrot_smoother_matches_reference():
    out = S.smooth_objrot(_sd("objrot."), G["obj_rot_in"], int(G["W"]))
    assert rel_err(out, G["obj_angles_out"]) < 1e-5
