"""demo.sh step 5 on the device (vistracker_b200/pipeline.py) against the composition of the pinned CPU restatements: PCA axes -> rotation
(oracle/geom_ref.py), object SmoothNet (oracle/smooth_ref.py), HVOP-Net in-filling (oracle/infill_ref.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import geom_ref as G
from oracle import infill_ref as I
from oracle import smooth_ref as SR
from vistracker_b200.synth import synthetic_infill_sequence, synthetic_infill_state_dict

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_object_rotation_stage_matches_oracle_composition():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from vistracker_b200.infill import CondMotionInfillAutoreg, ConditionalMInfiller
    from vistracker_b200.pipeline import object_rotation_stage, pack_neural
    from vistracker_b200.smooth import ObjrotSmoother
    dev = torch.device("cuda", 0)
    gi, gs = np.load(os.path.join(GOLD, "infill_small.npz")), np.load(os.path.join(GOLD, "smooth_small.npz"))
    opt = json.loads(str(gi["opt_json"]))
    sd_inf = synthetic_infill_state_dict(opt, seed=21)
    sd_obj = {k[7:]: torch.from_numpy(gs[k]) for k in gs.files if k.startswith("objrot.")}
    T = 260
    rng = np.random.default_rng(11)
    # a smooth object motion seen through noisy PCA-axis predictions of a box-like template; visibility drops in two spans
    src = torch.from_numpy(gi["pca_src"])
    t = np.arange(T)[:, None] / 40.0
    aa = np.array([[0.4, 1.0, -0.3]]) + 0.6 * np.sin(t * np.array([[0.8, 0.5, 1.1]]))
    Rt = SR.rot6d_to_rotmat(torch.from_numpy(SR.axis_to_rot6d_np(aa))).float()                       # true rotations
    pca = (src[None] @ Rt + 0.03 * torch.from_numpy(rng.standard_normal((T, 3, 3))).float()).contiguous()
    _, _, _, trans_obj, occ = synthetic_infill_sequence(T, seed=8, occluded=((60, 110), (170, 215)))
    poses = (0.3 * np.sin(t * rng.uniform(0.5, 2.0, (1, 156)) + rng.uniform(0, 6, (1, 156)))).astype(np.float32)
    trans = (np.array([[0.1, -0.2, 2.3]]) + 0.3 * np.sin(t * np.array([[0.7, 1.1, 0.4]]))).astype(np.float32)

    # ---- CPU: the reference's data flow, restated
    R0 = G.init_object_orientation(pca, src).float().numpy()
    angles_s = SR.smooth_objrot(sd_obj, R0.transpose(0, 2, 1))
    rot6d_obj = angles_s.transpose(1, 2)[:, :, :2].reshape(T, 6).numpy()
    smpl72 = SR.smplh_to_smpl_pose(poses)
    rot6d_smpl = SR.axis_to_rot6d_np(smpl72.reshape(-1, 3)).reshape(T, 144)
    ang_ref, trans_ref, _ = I.autoreg_infill(sd_inf, opt, rot6d_smpl, trans, rot6d_obj, trans_obj, occ, occ_thres=0.5)

    # ---- device
    net = ConditionalMInfiller(opt, device=dev).load_state_dict(sd_inf)
    neural = pack_neural(pca.to(dev), torch.zeros(T, 3, device=dev), torch.from_numpy(occ).to(dev))
    assert neural.shape == (T, 13)
    out = object_rotation_stage(neural, torch.from_numpy(poses).to(dev), torch.from_numpy(trans).to(dev), torch.from_numpy(trans_obj).to(dev),
                                src.to(dev), ObjrotSmoother(sd_obj, device=dev), CondMotionInfillAutoreg(net), occ_thres=0.5)
    assert out["infilled"]
    assert np.abs(out["obj_angles_smooth"].cpu().numpy() - angles_s.numpy()).max() < 1e-4              # fp32 SVD of noisy axes on the CPU side
    assert np.abs(out["obj_angles"].cpu().numpy() - ang_ref).max() < 3e-4
    assert np.array_equal(out["obj_trans"].cpu().numpy(), trans_ref)
    # nothing visible: HVOP-Net skips, the smoothed rotations pass through
    neural[:, 12] = 0.05
    out2 = object_rotation_stage(neural, torch.from_numpy(poses).to(dev), torch.from_numpy(trans).to(dev), torch.from_numpy(trans_obj).to(dev),
                                 src.to(dev), ObjrotSmoother(sd_obj, device=dev), CondMotionInfillAutoreg(net))
    assert not out2["infilled"] and torch.equal(out2["obj_angles"], out2["obj_angles_smooth"])
