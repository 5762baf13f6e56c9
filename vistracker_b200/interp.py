"""SLERP / LERP in-filling of occluded object poses: the non-learned baseline next to HVOP-Net (interp/interpolate_recon.py:24-171,
interp/lib/quaternions.py:38-68).  A few hundred quaternions per sequence -- host arithmetic like the reference's, no kernel: the
sequence-global stages hand over [T, 3, 3] rotations, this returns them with the occluded spans replaced.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch


def compute_missing_inds(mask: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """mask [T] (1 = invisible) -> (end_inds, start_inds) of the spans to interpolate (interpolate_recon.py:66-79)."""
    mask = np.asarray(mask, dtype=float)
    diff = mask - np.concatenate([mask[0:1], mask[:-1]])
    return np.where(diff == -1)[0], np.where(diff == 1)[0]


def slerp(q0: torch.Tensor, q1: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """interp/lib/quaternions.py:38-68, shapes (B, J, 4), (B, J, 4), (B, T, J, 1) -> (B, T, J, 4).  As there, the end quaternion is flipped to
    the near hemisphere but the angle is taken from the UNFLIPPED dot product."""
    assert q0.shape == q1.shape, f"shape does not match: q0={q0.shape}, q1={q1.shape}"
    q0, q1 = q0.unsqueeze(1), q1.unsqueeze(1)
    cos_half = torch.sum(q0 * q1, dim=-1, keepdim=True)
    q1 = torch.where(cos_half > 0, q1, -q1)
    half = torch.acos(cos_half)
    sin_half = torch.sqrt(1.0 - cos_half * cos_half)
    qt = torch.sin((1 - t) * half) / sin_half * q0 + torch.sin(t * half) / sin_half * q1
    return torch.where(torch.abs(cos_half) >= 1.0, q0, qt)


def _spans(start_inds, end_inds):
    """(clip_start, clip_end, next_start or None) per span, with the reference's bookkeeping (interpolate_recon.py:97-118)."""
    for i, (start, end) in enumerate(zip(start_inds, end_inds)):
        last = start == start_inds[-1] or end == end_inds[-1]
        yield int(start) - 1, int(end), None if last else int(start_inds[i + 1]) - 1


def interp_lerp(end_inds, transl: np.ndarray, start_inds) -> np.ndarray:
    """Linear interpolation of translations [T, 3] over the spans (interpolate_recon.py:81-121)."""
    assert start_inds[0] >= 1
    out = [transl[:start_inds[0] - 1]]
    for c0, c1, nxt in _spans(start_inds, end_inds):
        n = c1 - c0
        a, b = transl[c0:c0 + 1], transl[c1:c1 + 1]
        times = np.arange(1, n) / n
        out += [a, np.expand_dims(times, -1).repeat(3, -1) * (b - a) + a, transl[c1:] if nxt is None else transl[c1:nxt]]
    return np.concatenate(out, 0)


def interp_slerp(end_inds, rot_q: np.ndarray, start_inds) -> np.ndarray:
    """SLERP of quaternions [T, 4] over the spans (interpolate_recon.py:123-165)."""
    assert start_inds[0] >= 1
    out = [rot_q[:start_inds[0] - 1]]
    for c0, c1, nxt in _spans(start_inds, end_inds):
        n = c1 - c0
        a, b = rot_q[c0:c0 + 1], rot_q[c1:c1 + 1]
        times = np.arange(1, n) / n
        q = slerp(torch.from_numpy(a).unsqueeze(0), torch.from_numpy(b).unsqueeze(0), torch.from_numpy(times).unsqueeze(0).unsqueeze(-1).unsqueeze(-1))
        out += [a, q[0, :, 0, :].cpu().numpy(), rot_q[c1:] if nxt is None else rot_q[c1:nxt]]
    return np.concatenate(out, 0)


def interpolate_object_rotations(obj_angles, occ_ratios, thres: float = 0.3, obj_trans: Optional[np.ndarray] = None):
    """``BaseInterpolator.interp_seq`` without the joblib IO (interpolate_recon.py:29-64, 167-170): ``obj_angles`` [T,3,3] as stored (= R^T),
    ``occ_ratios`` [T] visible fraction.  Frames with ``occ_ratios < thres`` are replaced by SLERP between the visible neighbours; a leading
    occluded span is left alone (the reference only warns).  Returns ``obj_angles`` in the same layout (and the LERP-ed translation when
    ``obj_trans`` is given)."""
    from scipy.spatial.transform import Rotation
    as_np = lambda a: a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    ang, occ = as_np(obj_angles), as_np(occ_ratios)
    rot_q = Rotation.from_matrix(ang.transpose(0, 2, 1)).as_quat()
    end_inds, start_inds = compute_missing_inds((occ < thres).astype(float))
    if len(start_inds) == 0:
        q, tr = rot_q, (None if obj_trans is None else as_np(obj_trans))
    else:
        if end_inds[0] < start_inds[0]:
            end_inds = end_inds[1:]
        q = interp_slerp(end_inds, rot_q, start_inds)
        tr = None if obj_trans is None else interp_lerp(end_inds, as_np(obj_trans), start_inds)
        assert len(q) == len(rot_q), f"GT={len(rot_q)}, interpolate={len(q)}"
    out = Rotation.from_quat(q).as_matrix().transpose(0, 2, 1)
    return out if obj_trans is None else (out, tr)
