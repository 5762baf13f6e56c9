"""Sequence evaluation on B200: the numeric core of ``VideoPackedEvaluator.eva_seq``
(recon/eval/evalvideo_packed.py:100-147) -- per-window Procrustes alignment of the reconstructed SMPL + object vertices to the ground
truth (``compute_transform``, recon/eval/pose_utils.py:153-198), then per frame the bidirectional Chamfer distance between surface
samples (recon/eval/evaluate.py:126-160, chamfer_distance.py:10-52) and the vertex-to-vertex error (evaluate.py:172-174), in cm.

The reference builds psbody meshes, samples with trimesh, queries sklearn kd-trees and runs all of it frame by frame on the CPU; here the
sequence is a handful of device tensors and three kernels (vt_procrustes, vt_similarity_apply, vt_nn_dist).  Loading the packed
reconstruction / ground-truth files and the SMPL forward that produces the vertices stay with the caller.
"""
from __future__ import annotations

from typing import Optional

import torch

from .geom import apply_similarity, eval_chamfer_distance, procrustes_transform

UNIT_CVT = 100.0          # metres -> centimetres (recon/eval/evaluate.py: self.unit_cvt)


def sample_surface(verts: torch.Tensor, faces: torch.Tensor, n: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Area-weighted uniform surface samples, [B, V, 3] x [F, 3] -> [B, n, 3] (the role of ``trimesh.Trimesh.sample`` in
    ReconEvaluator.surface_sampling; trimesh draws from numpy's global generator, so the samples themselves are not reproducible there)."""
    tri = verts[:, faces.long()]                                              # [B, F, 3, 3]
    area = torch.linalg.cross(tri[:, :, 1] - tri[:, :, 0], tri[:, :, 2] - tri[:, :, 0], dim=-1).norm(dim=-1)
    fidx = torch.multinomial(area, n, replacement=True, generator=generator)  # [B, n]
    uv = torch.rand(verts.shape[0], n, 2, device=verts.device, generator=generator)
    flip = uv.sum(-1) > 1
    uv = torch.where(flip[..., None], 1 - uv, uv)
    t = torch.gather(tri, 1, fidx[:, :, None, None].expand(-1, -1, 3, 3))
    return t[:, :, 0] + uv[..., :1] * (t[:, :, 1] - t[:, :, 0]) + uv[..., 1:] * (t[:, :, 2] - t[:, :, 0])


def evaluate_sequence(sverts_recon, overts_recon, sverts_gt, overts_gt, smpl_faces=None, obj_faces=None, window: int = 300,
                      recon_exist=None, sample_num: Optional[int] = 10000, smpl_only: bool = False, generator=None):
    """Per-frame errors [n_valid, 4] = (Chamfer SMPL, Chamfer object, v2v SMPL, v2v object) in cm for a sequence of reconstructed and
    ground-truth vertices ([T, Vs, 3] / [T, Vo, 3] CUDA tensors), with one similarity alignment per window of ``window`` frames exactly
    as the reference schedules it (a new alignment on the first frame and whenever the running frame count is a multiple of the window;
    window <= 0: no alignment).  ``sample_num=None`` evaluates the Chamfer distance on the vertices instead of surface samples.
    Returns (errors, kept frame indices, list of (first frame, R, t, scale))."""
    T = sverts_gt.shape[0]
    exist = torch.ones(T, dtype=torch.bool) if recon_exist is None else torch.as_tensor(recon_exist).bool().cpu()
    do_align = window > 0
    count, cur = 0, None
    transforms, keep, aligned_s, aligned_o = [], [], [], []
    for i in range(T):
        count += 1
        if do_align:
            if cur is None or count % window == 0:
                idx = torch.arange(i, min(T, i + window))
                idx = idx[exist[idx]]
                if len(idx) == 0:
                    continue
                if smpl_only:
                    src, dst = sverts_recon[idx].reshape(-1, 3), sverts_gt[idx].reshape(-1, 3)
                else:
                    src = torch.cat([sverts_recon[idx].reshape(-1, 3), overts_recon[idx].reshape(-1, 3)], 0)
                    dst = torch.cat([sverts_gt[idx].reshape(-1, 3), overts_gt[idx].reshape(-1, 3)], 0)
                cur = procrustes_transform(src, dst)
                transforms.append((i,) + cur)
        if not bool(exist[i]):
            continue
        keep.append(i)
        if do_align:
            aligned_s.append(apply_similarity(sverts_recon[i], *cur))
            aligned_o.append(apply_similarity(overts_recon[i], *cur))
        else:
            aligned_s.append(sverts_recon[i]); aligned_o.append(overts_recon[i])
    if not keep:
        return torch.zeros(0, 4, device=sverts_gt.device), keep, transforms
    k = torch.as_tensor(keep, device=sverts_gt.device)
    rs, ro, gs, go = torch.stack(aligned_s), torch.stack(aligned_o), sverts_gt[k], overts_gt[k]
    if sample_num is None:
        ps_r, po_r, ps_g, po_g = rs, ro, gs, go
    else:
        ps_g, ps_r = sample_surface(gs, smpl_faces, sample_num, generator), sample_surface(rs, smpl_faces, sample_num, generator)
        po_g, po_r = sample_surface(go, obj_faces, sample_num, generator), sample_surface(ro, obj_faces, sample_num, generator)
    errs = torch.stack([eval_chamfer_distance(ps_g, ps_r), eval_chamfer_distance(po_g, po_r),
                        (gs - rs).norm(dim=-1).mean(-1), (go - ro).norm(dim=-1).mean(-1)], 1) * UNIT_CVT
    return errs, keep, transforms


def acceleration_errors(verts_recon: torch.Tensor, verts_gt: torch.Tensor, transforms, window: int, recon_exist=None) -> torch.Tensor:
    """The 'smpl-acc' / 'obj-acc' column of ``VideoPackedEvaluator.eva_seq`` (recon/eval/evalvideo_packed.py:148-160 over
    ``compute_accel_err``, recon/eval/evaluate_video.py:138-157) for one of the two meshes: per kept frame, the mean norm (cm) of the difference
    of the second temporal differences of aligned reconstruction and ground truth, computed per flushed group -- a group ends when the
    running frame count is a multiple of ``window`` or at the last frame, provided that frame has a reconstruction -- and repeated for the
    frames of the group (NaN for groups of fewer than three frames).  ``transforms``: the (first frame, R, t, scale) list that
    ``evaluate_sequence`` returns; frames use the latest alignment computed at or before them.  Plain tensor arithmetic on the device the
    vertices live on."""
    L = verts_gt.shape[0]
    exist = torch.ones(L, dtype=torch.bool) if recon_exist is None else torch.as_tensor(recon_exist).bool().cpu()
    starts = [int(t[0]) for t in transforms]
    out, group_r, group_g, count, ti = [], [], [], 0, -1
    for i in range(L):
        count += 1
        realign = (ti < 0 or count % window == 0)
        if realign:
            if i in starts:
                ti = starts.index(i)
            elif not bool(exist[i:min(L, i + window)].any()):
                continue                                                           # no alignment possible: the reference skips the frame entirely
        if not bool(exist[i]):
            continue
        _, R, t, s = transforms[ti]
        R, t = torch.as_tensor(R, dtype=verts_recon.dtype, device=verts_recon.device), torch.as_tensor(t, dtype=verts_recon.dtype, device=verts_recon.device)
        group_r.append(float(s) * verts_recon[i] @ R.T + t.reshape(1, 3))
        group_g.append(verts_gt[i])
        if count % window == 0 or i == L - 1:
            g, r = torch.stack(group_g), torch.stack(group_r)
            d = (g[:-2] - 2 * g[1:-1] + g[2:]) - (r[:-2] - 2 * r[1:-1] + r[2:])
            val = d.norm(dim=2).mean() * UNIT_CVT if d.numel() else torch.tensor(float("nan"), device=verts_gt.device, dtype=verts_gt.dtype)
            out += [val] * len(group_g)
            group_r, group_g = [], []
    return torch.stack(out) if out else torch.zeros(0, device=verts_gt.device, dtype=verts_gt.dtype)
