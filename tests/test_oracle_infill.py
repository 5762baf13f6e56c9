"""oracle/infill_ref.py (HVOP-Net restatement) against the reference's own outputs (tests/golden/infill_small.npz) -- CPU only."""
import json
import os

import numpy as np
import torch

from oracle import infill_ref as R
from vistracker_b200.synth import infill_spec, synthetic_infill_sequence, synthetic_infill_state_dict

GOLD = os.path.join(os.path.dirname(__file__), "golden", "infill_small.npz")


def _load():
    g = np.load(GOLD)
    opt = json.loads(str(g["opt_json"]))
    return g, opt, synthetic_infill_state_dict(opt, seed=21)


def test_forward_matches_reference_module():
    g, opt, sd = _load()
    for tag in ("a", "b"):
        pred = R.cond_infiller_forward(sd, opt, g[f"{tag}_data_smpl"], g[f"{tag}_mask_smpl"], g[f"{tag}_data_obj"], g[f"{tag}_mask_obj"])
        err = np.abs(pred.numpy() - g[f"{tag}_pred"]).max() / np.abs(g[f"{tag}_pred"]).max()
        assert err < 2e-5, (tag, err)


def test_autoregressive_loop_matches_reference_test_loop():
    g, opt, sd = _load()
    L = int(g["seq_L"])
    seq = synthetic_infill_sequence(L, seed=5)
    for a, k in zip(seq, ("seq_rot6d_smpl", "seq_trans_smpl", "seq_rot6d_obj", "seq_trans_obj", "seq_occ")):
        assert np.array_equal(a, g[k])                                            # the generator of the inputs is deterministic
    angles, trans, rot6d = R.autoreg_infill(sd, opt, *seq, occ_thres=0.5)
    assert np.abs(angles - g["seq_obj_angles"]).max() < 5e-5
    assert np.array_equal(trans, g["seq_obj_trans"])
    assert np.abs(np.linalg.det(angles) - 1).max() < 1e-6 and bool((g["seq_obj_scales"] == 1).all())


def test_no_seed_frames_is_skipped():
    g, opt, sd = _load()
    seq = list(synthetic_infill_sequence(200, seed=1))
    seq[4] = np.full(200, 0.1, np.float32)                                        # everything occluded: fewer than 30 visible frames in clip 0
    assert R.autoreg_infill(sd, opt, *seq) is None


def test_position_embedding_shapes_and_spec():
    assert R.position_embedding(180, 160).shape == (180, 160) and R.position_embedding(7, 33).shape == (7, 33)
    assert float(R.position_embedding(1, 8).abs().max()) <= 1.0
    g, opt, sd = _load()
    assert [k for k, _, _ in infill_spec(opt)] == list(sd.keys())


def test_pca_orientation_matches_reference_pcautil():
    from oracle import geom_ref as G
    g = np.load(GOLD)
    R_ref = G.init_object_orientation(torch.from_numpy(g["pca_tgt"]), torch.from_numpy(g["pca_src"]))
    assert np.abs(R_ref.numpy() - g["pca_R"]).max() < 2e-5
    assert np.abs(np.linalg.det(g["pca_R"]) - 1).max() < 1e-5
