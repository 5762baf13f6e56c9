"""CPU restatement of the SMPL-T keypoint pre-fit objective and optimiser schedule (TEST INFRASTRUCTURE).

Follows preprocess/fit_SMPLH_30fps.py:26-66 (joint weights, loss weights), :153-200 (compute_loss, priors, second-difference
smoothness terms), preprocess/fit_SMPLH_kpts.py:67-75 (sum_dict), :114-190 (fit_one_batch loop, the two Adam phases),
:306-310 (project_points), lib_smpl/th_smpl_prior.py:25-39, lib_smpl/th_hand_prior.py:46-72 (incl. the [1, 2B, 45]
concatenation that makes the hand term  sum / 45  rather than a batch mean).

PINNED against the unmodified reference methods (tests/golden/fit_smplt_small.npz, make_golden.py).
"""
from __future__ import annotations

import numpy as np
import torch

from .smpl_ref import landmarks, smpl_forward

FX, FY, CX, CY = 979.7844, 979.840, 1018.952, 779.486          # fit_SMPLH_kpts.py:46-47
JOINT_WEIGHTS = np.array([1, 1, 1, 10, 10, 10, 10, 10, 10, 10, 10, 10, 5, 5, 5, 5, 5, 5, 10, 10, 10, 1, 1, 1, 1, 1, 1, 10, 10, 10,
                          1, 1, 1, 1, 1, 1, 5, 10, 10, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1],
                         np.float64)                              # fit_SMPLH_30fps.py:26-51
LOSS_W = {"pose": 1e-5, "hand": 1e-5, "kpts": 0.3 ** 2, "temp": 30.0 ** 2, "ptemp": 5.0 ** 2, "pinit": 30.0 ** 2}


def compute_loss(model, reg, priors, pose, betas, trans, kpts, pose_init):
    """SMPLHFitter30fps.compute_loss.  reg = (row, col, val, shape) of the body-25 regressor; priors = dict of
    body/lh/rh mean + precision.  Returns the dict of UNWEIGHTED terms."""
    dt = pose.dtype
    verts, _, _, _ = smpl_forward(model, pose, betas, trans, torch.zeros(pose.shape[0], verts_count(model), 3, dtype=dt))
    J = landmarks(torch.stack([torch.as_tensor(reg[0]).long(), torch.as_tensor(reg[1]).long()]), torch.as_tensor(reg[2]), tuple(reg[3]), verts)
    px = J[:, :, 0:1] * FX / J[:, :, 2:3] + CX
    py = J[:, :, 1:2] * FY / J[:, :, 2:3] + CY
    proj = torch.cat([px, py], -1)
    loss = {"kpts": ((proj - kpts[:, :, :2]) ** 2 * kpts[:, :, 2:3]).mean()}
    loss["temp"] = (((verts[1:-1] - verts[:-2]) - (verts[2:] - verts[1:-1])) ** 2).mean()
    v1, v2 = pose[1:-1, :66] - pose[:-2, :66], pose[2:, :66] - pose[1:-1, :66]
    loss["ptemp"] = (((v1 - v2) ** 2) * torch.as_tensor(JOINT_WEIGHTS)[None]).mean()       # float64 weights promote, as in the reference
    t = (pose[:, 3:66] - torch.as_tensor(priors["body_prior_mean"], dtype=torch.float32).to(dt)) @ \
        torch.as_tensor(priors["body_prior_precision"].astype(np.float32)).to(dt)
    loss["pose"] = (t * t).sum(1).mean()
    hm = torch.cat([torch.as_tensor(priors["lh_prior_mean"], dtype=torch.float32), torch.as_tensor(priors["rh_prior_mean"], dtype=torch.float32)]).to(dt)
    th = pose[:, 66:] - hm
    lh = th[:, :45] @ torch.as_tensor(priors["lh_prior_precision"], dtype=torch.float32).to(dt)
    rh = th[:, 45:] @ torch.as_tensor(priors["rh_prior_precision"], dtype=torch.float32).to(dt)
    loss["hand"] = ((lh * lh).sum() + (rh * rh).sum()) / 45.0
    loss["pinit"] = ((pose_init[:, 3:66] - pose[:, 3:66]) ** 2).mean()
    return loss


def verts_count(model):
    return model["th_v_template"].shape[1]


def total_loss(loss, decay):
    """BaseFitter.sum_dict with the 30fps weights: sum_k w_k * loss_k / (1 + decay)."""
    return sum(LOSS_W[k] * v / (1 + decay) for k, v in loss.items())


def fit(model, reg, priors, pose0, betas0, trans0, kpts, n_outer, steps_per_iter=10, iter_for_global=8, record=()):
    """fit_one_batch without IO and without early stop: Adam(lr .01) on [trans, global_pose, top_betas] for the first
    `iter_for_global` outer iterations, then a NEW Adam(lr .001) on [trans, global_pose, body_pose, top_betas, other_betas]."""
    gp, bp, hp = (pose0[:, :3].clone().requires_grad_(True), pose0[:, 3:66].clone().requires_grad_(True), pose0[:, 66:].clone())
    tb, ob = betas0[:, :2].clone().requires_grad_(True), betas0[:, 2:].clone().requires_grad_(True)
    tr = trans0.clone().requires_grad_(True)
    pose_init = pose0.clone()
    opt = torch.optim.Adam([tr, gp, tb], lr=0.01)
    losses, snaps, step = [], {}, 0
    for it in range(n_outer):
        if it == iter_for_global:
            opt = torch.optim.Adam([tr, gp, bp, tb, ob], lr=0.001)
        for _ in range(steps_per_iter):
            opt.zero_grad()
            pose, betas = torch.cat([gp, bp, hp], 1), torch.cat([tb, ob], 1)
            ld = compute_loss(model, reg, priors, pose, betas, tr, kpts, pose_init)
            loss = total_loss(ld, it // 3)
            loss.backward()
            opt.step()
            losses.append(float(loss))
            step += 1
            if step in record:
                snaps[step] = (torch.cat([gp, bp, hp], 1).detach().clone(), torch.cat([tb, ob], 1).detach().clone(), tr.detach().clone())
    return torch.cat([gp, bp, hp], 1).detach(), torch.cat([tb, ob], 1).detach(), tr.detach(), losses, snaps
