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


def compute_transform(S1, S2):
    """recon/eval/pose_utils.py:153-198 restated (float64 numpy SVD): (R, t, scale) with scale * R @ p + t taking S1 [N, 3] onto S2."""
    import numpy as np
    S1, S2 = np.asarray(S1, np.float64).T, np.asarray(S2, np.float64).T
    mu1, mu2 = S1.mean(axis=1, keepdims=True), S2.mean(axis=1, keepdims=True)
    X1, X2 = S1 - mu1, S2 - mu2
    var1 = np.sum(X1 ** 2)
    K = X1.dot(X2.T)
    U, s, Vh = np.linalg.svd(K)
    V = Vh.T
    Z = np.eye(3)
    Z[-1, -1] *= np.sign(np.linalg.det(U.dot(V.T)))
    R = V.dot(Z.dot(U.T))
    scale = np.trace(R.dot(K)) / var1
    t = mu2 - scale * (R.dot(mu1))
    return R, t[:, 0], scale


def evaluate_sequence(sverts_recon, overts_recon, sverts_gt, overts_gt, window, recon_exist=None, smpl_only=False, with_accel=False,
                      return_transforms=False):
    """The alignment-window loop of VideoPackedEvaluator.eva_seq (recon/eval/evalvideo_packed.py:100-163) restated on plain arrays, with
    the Chamfer distance evaluated on the vertices (the reference's trimesh surface samples are unseeded random draws): rows of
    (Chamfer SMPL, Chamfer object, v2v SMPL, v2v object) in cm for every frame that has a reconstruction; ``with_accel`` appends the two
    acceleration-error columns (:148-160 over evaluate_video.py:138-157: one value per flushed group of frames, repeated).  Pinned by
    tests/golden/eval_seq.npz -- the reference's own eva_seq run on in-memory arrays."""
    import numpy as np
    L = len(sverts_gt)
    exist = np.ones(L, bool) if recon_exist is None else np.asarray(recon_exist, bool)
    count, arot, out, transforms = 0, None, [], []
    col = [[], [], [], []]                           # aligned SMPL recon, SMPL gt, aligned object recon, object gt since the last flush
    acc_s, acc_o = [], []

    def accel(gt, rc):
        gt, rc = np.stack(gt, 0), np.stack(rc, 0)
        d = (gt[:-2] - 2 * gt[1:-1] + gt[2:]) - (rc[:-2] - 2 * rc[1:-1] + rc[2:])
        n = np.linalg.norm(d, axis=2)
        return float(n.mean() * 100) if n.size else float("nan")

    for i in range(L):
        count += 1
        rec = [np.asarray(sverts_recon[i], np.float64), np.asarray(overts_recon[i], np.float64)]
        if window > 0:
            if arot is None or count % window == 0:
                idx = np.arange(i, min(L, i + window))
                idx = idx[exist[idx]]
                if len(idx) == 0:
                    continue
                if smpl_only:
                    g, r = np.concatenate(sverts_gt[idx], 0), np.concatenate(sverts_recon[idx], 0)
                else:
                    g = np.concatenate([np.concatenate(x[idx], 0) for x in (sverts_gt, overts_gt)], 0)
                    r = np.concatenate([np.concatenate(x[idx], 0) for x in (sverts_recon, overts_recon)], 0)
                arot, atrans, ascale = compute_transform(r, g)
                transforms.append((i, arot, atrans, ascale))
            rec = [(ascale * arot.dot(m.T) + atrans[:, None]).T for m in rec]
        if not exist[i]:
            continue
        gt = [np.asarray(sverts_gt[i], np.float64), np.asarray(overts_gt[i], np.float64)]
        row = [eval_chamfer(g, r) * 100.0 for g, r in zip(gt, rec)] + [np.sqrt(((g - r) ** 2).sum(-1)).mean() * 100.0 for g, r in zip(gt, rec)]
        out.append(row)
        if with_accel:
            col[0].append(rec[0]); col[1].append(gt[0]); col[2].append(rec[1]); col[3].append(gt[1])
            if count % window == 0 or i == L - 1:
                n = len(col[1])
                acc_s += [accel(col[1], col[0])] * n
                acc_o += [accel(col[3], col[2])] * n
                col = [[], [], [], []]
    out = np.asarray(out)
    if with_accel:
        out = np.concatenate([out, np.asarray(acc_s)[:, None], np.asarray(acc_o)[:, None]], 1)
    return (out, transforms) if return_transforms else out


def init_object_orientation(tgt_axis: torch.Tensor, src_axis: torch.Tensor, noise: torch.Tensor = None) -> torch.Tensor:
    """recon/recon_fit_base.py:202-225 (inverse = (S^T S)^-1 S^T, then decopose_axis) and recon/pca_util.py:59-72 (noise None)."""
    S, T = src_axis.double(), tgt_axis.double()
    if S.dim() == 2:
        S = S[None].expand(T.shape[0], 3, 3)
    pseudo = torch.linalg.inv(S.transpose(1, 2) @ S) @ S.transpose(1, 2)
    rot = pseudo @ T
    if noise is not None:
        rot = rot + 1e-4 * noise.double()
    return project_so3(rot)
