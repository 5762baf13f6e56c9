"""CPU restatements of the small geometric operators of the joint optimisation (TEST INFRASTRUCTURE).

* ``project_so3`` / ``decopose_axis`` / ``transform_obj_verts`` follow recon/recon_fit_base.py:178-199,455-469 literally
  (torch.svd -> det -> flip the last row of V^T) -- the reference code is in-tree, so this part is a 1:1 restatement.
* ``chamfer_ragged`` restates ``pytorch3d.loss.chamfer_distance(Pointclouds, Pointclouds)`` with default arguments
  (call site recon/recon_fit_trivis_full.py:452-456).  pytorch3d is an UNPINNED, un-vendored dependency of the reference
  (requirements.txt:22) and is not installable here: PARITY UNPINNED for this function -- it follows the published
  semantics (squared L2 nearest neighbour, mean over the valid points of each cloud, mean over the batch, sum of both
  directions).
"""
from __future__ import annotations

from typing import List

import torch


def project_so3(mat: torch.Tensor) -> torch.Tensor:
    u, s, v = torch.svd(mat)
    vt = v.transpose(1, 2)
    det = torch.det(torch.matmul(u, vt)).view(-1, 1, 1)
    vt = torch.cat((vt[:, :2, :], vt[:, -1:, :] * det), 1)
    return torch.matmul(u, vt)


def transform_obj_verts(verts, obj_R, obj_t, obj_s):
    """(P R + t) * s -- row-vector convention, scale after rotation and translation (recon_fit_base.py:455-459)."""
    return (torch.bmm(verts, obj_R) + obj_t.unsqueeze(1)) * obj_s.unsqueeze(1).unsqueeze(1)


def chamfer_ragged(xs: List[torch.Tensor], ys: List[torch.Tensor]) -> torch.Tensor:
    n = len(xs)
    total = xs[0].new_zeros(())
    for x, y in zip(xs, ys):
        d = ((x[:, None, :] - y[None, :, :]) ** 2).sum(-1)
        total = total + d.min(1).values.mean() + d.min(0).values.mean()
    return total / n


def eval_chamfer(x, y, direction="bi"):
    """recon/eval/chamfer_distance.py:10-52 restated with a brute-force float64 nearest neighbour (the reference uses an exact sklearn
    kd-tree on the same metric, so the two agree to rounding): mean Euclidean NN distance, both directions summed for 'bi'.
    Pinned by tests/golden/eval_chamfer.npz (the reference function itself, run by tests/golden/make_golden.py --only eval)."""
    import numpy as np
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    d = np.sqrt(((x[:, None, :] - y[None, :, :]) ** 2).sum(-1))
    x_to_y, y_to_x = d.min(1).mean(), d.min(0).mean()
    return {"bi": x_to_y + y_to_x, "x_to_y": x_to_y, "y_to_x": y_to_x}[direction]
