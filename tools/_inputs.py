"""Inputs shared by the tools (timing and integration runs): the reference-derived asset arrays and a seeded SMPL-T fitting problem.
Unlike tests/fit_problem.py this module does not touch the CPU restatements under oracle/ (test infrastructure): the 2-D key points are
projections of a plausible joint cloud around the body translation, which is all a timing run needs."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_assets():
    """(dict of arrays from tests/golden/assets.npz -- priors, part labels, body-25 regressor of the reference's assets/ --, COO regressor tuple)."""
    a = dict(np.load(os.path.join(ROOT, "tests", "golden", "assets.npz")))
    return a, (a["body25_row"], a["body25_col"], a["body25_val"], a["body25_shape"])


def synthetic_fit_problem(frames: int, seed: int):
    """(synthetic SMPL-H model, kpts [T,25,3] (x, y in 2048x1536 pixels, confidence), pose0 [T,156], betas0 [T,10], trans0 [T,3])."""
    from vistracker_b200.synth_smpl import synthetic_motion, synthetic_smplh
    model = synthetic_smplh(seed=3)
    pose, betas, trans = synthetic_motion(frames, seed=seed)
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    off = torch.from_numpy((rng.standard_normal((1, 25, 3)) * np.array([0.25, 0.45, 0.12])).astype(np.float32))
    J = trans[:, None, :] + off                                                      # a rigid joint cloud riding on the translation
    k2d = torch.stack([J[..., 0] * 979.7844 / J[..., 2] + 1018.952, J[..., 1] * 979.840 / J[..., 2] + 779.486], -1)
    k2d = k2d + torch.from_numpy(rng.standard_normal(tuple(k2d.shape)).astype(np.float32)) * 2.0
    conf = torch.from_numpy(rng.uniform(0.3, 1.0, (frames, 25, 1)).astype(np.float32))
    conf[torch.from_numpy(rng.random((frames, 25, 1)) < 0.1)] = 0.0
    pose0 = pose.clone()
    pose0[:, :66] += torch.from_numpy(rng.standard_normal((frames, 66)).astype(np.float32)) * 0.1
    betas0 = torch.zeros(frames, 10); betas0[:, 0] = 2.2
    trans0 = trans + torch.from_numpy(rng.standard_normal((frames, 3)).astype(np.float32)) * 0.05
    return model, torch.cat([k2d, conf], -1), pose0, betas0, trans0
