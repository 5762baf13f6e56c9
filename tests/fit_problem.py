"""Seeded SMPL-T fitting problem shared by the oracle test, the GPU test and make_golden.py (SURVEY.md 8(d) C3 style)."""
import os

import numpy as np
import torch

from oracle.smpl_ref import landmarks, smpl_forward
from vistracker_b200.synth_smpl import synthetic_motion, synthetic_smplh

HERE = os.path.dirname(os.path.abspath(__file__))


def load_assets():
    a = dict(np.load(os.path.join(HERE, "golden", "assets.npz")))
    reg = (a["body25_row"], a["body25_col"], a["body25_val"], a["body25_shape"])
    return a, reg


def synthetic_fit_problem(frames: int, seed: int):
    a, reg = load_assets()
    model = synthetic_smplh(seed=3)
    pose, betas, trans = synthetic_motion(frames, seed=seed)
    with torch.no_grad():
        verts = smpl_forward(model, pose, betas, trans)[0]
        J = landmarks(torch.stack([torch.as_tensor(reg[0]).long(), torch.as_tensor(reg[1]).long()]), torch.as_tensor(reg[2]),
                      tuple(reg[3]), verts)
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    k2d = torch.stack([J[..., 0] * 979.7844 / J[..., 2] + 1018.952, J[..., 1] * 979.840 / J[..., 2] + 779.486], -1)
    k2d = k2d + torch.from_numpy(rng.standard_normal(tuple(k2d.shape)).astype(np.float32)) * 2.0
    conf = torch.from_numpy(rng.uniform(0.3, 1.0, (frames, 25, 1)).astype(np.float32))
    conf[torch.from_numpy(rng.random((frames, 25, 1)) < 0.1)] = 0.0
    kpts = torch.cat([k2d, conf], -1)
    pose0 = pose.clone()
    pose0[:, :66] += torch.from_numpy(rng.standard_normal((frames, 66)).astype(np.float32)) * 0.1
    betas0 = torch.zeros(frames, 10); betas0[:, 0] = 2.2
    trans0 = trans + torch.from_numpy(rng.standard_normal((frames, 3)).astype(np.float32)) * 0.05
    return model, kpts, pose0, betas0, trans0
