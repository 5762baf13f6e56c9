"""CPU restatement of the SmoothNet stage -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows, line by line, the reference's
  * ``SmoothNet.forward`` / ``SmoothNetResBlock.forward`` (smoothnet/models/smoothnet.py:27-38,125-141) in eval mode and
    ``SmoothNetSMPL.forward`` (smoothnet/models/smoothnet_smpl.py:26-47: pose net on 144 channels, translation net on 3, betas pass),
  * ``SmootherBase.seq2batches`` (smoothnet/smooth_base.py:45-73) and ``clips2seq_fast`` / ``slide_window_to_sequence``
    (smoothnet/utils/utils.py:63-103) for window step 1,
  * ``SMPLTSmoother.preprocess_input / post_processing`` (smoothnet/smooth_smplt.py:27-101) and
    ``ObjrotSmoother.preprocess_input / post_processing`` (smoothnet/smooth_objrot.py:73-121),
  * the rotation conversions of smoothnet/utils/geometry_utils.py (numpy_axis_to_rot6D :285-346, rot6d_to_rotmat :63-77,
    rotation_matrix_to_angle_axis :93-119, rotation_matrix_to_quaternion :169-247, quaternion_to_angle_axis :122-166).
Pinned by tests/golden/smooth_small.npz, produced by tests/golden/make_golden.py from the reference's own classes and functions.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ network
def smoothnet_forward(sd, prefix: str, x: torch.Tensor) -> torch.Tensor:
    """x [N, C, T] -> [N, C, T]; sd holds `<prefix>encoder.0.*`, `<prefix>res_blocks.<i>.linear{1,2}.*`, `<prefix>decoder.*`."""
    x = x.to(torch.float32)
    x = F.leaky_relu(F.linear(x, sd[prefix + "encoder.0.weight"], sd[prefix + "encoder.0.bias"]), 0.1)
    i = 0
    while f"{prefix}res_blocks.{i}.linear1.weight" in sd:
        ident = x
        y = F.leaky_relu(F.linear(x, sd[f"{prefix}res_blocks.{i}.linear1.weight"], sd[f"{prefix}res_blocks.{i}.linear1.bias"]), 0.2)
        y = F.leaky_relu(F.linear(y, sd[f"{prefix}res_blocks.{i}.linear2.weight"], sd[f"{prefix}res_blocks.{i}.linear2.bias"]), 0.2)
        x = y + ident
        i += 1
    return F.linear(x, sd[prefix + "decoder.weight"], sd[prefix + "decoder.bias"])


def seq2batches(seq: torch.Tensor, window: int) -> torch.Tensor:
    """[L, D] -> [L - window + 1, window, D] (step 1)."""
    return torch.stack([seq[i:i + window].clone() for i in range(seq.shape[0] - window + 1)], 0)


def slide_window_mean(clips: torch.Tensor) -> torch.Tensor:
    """clips2seq_fast for step 1: frame l is the mean over the windows that contain it.  [B, T, D] -> [B + T - 1, D]."""
    B, T, D = clips.shape
    L = B + T - 1
    out = torch.zeros(L, D, dtype=clips.dtype)
    cnt = torch.zeros(L, dtype=clips.dtype)
    for b in range(B):
        out[b:b + T] += clips[b]
        cnt[b:b + T] += 1
    return out / cnt[:, None]


# ------------------------------------------------------------------------------------------------ rotations
def axis_to_rot6d_np(axis: np.ndarray) -> np.ndarray:
    """numpy_axis_to_rot6D: [n, 3] -> [n, 6] (first two COLUMNS of R, row-major over the 3x2 block), in the dtype of the input."""
    l = np.linalg.norm(axis + 1e-8, ord=2, axis=1)
    angle = np.expand_dims(l, -1)
    normalized = axis / angle
    angle = angle * 0.5
    quat = np.concatenate((np.cos(angle), np.sin(angle) * normalized), axis=1)
    q = quat / np.linalg.norm(quat + 1e-8, ord=2, axis=1, keepdims=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    w2, x2, y2, z2 = w ** 2, x ** 2, y ** 2, z ** 2
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    R = np.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz, 2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                  2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], axis=1).reshape(-1, 3, 3)
    return R[:, :, :2].reshape(-1, 6)


def rot6d_to_rotmat(x: torch.Tensor) -> torch.Tensor:
    x = x.reshape(-1, 3, 2)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = F.normalize(a1)
    b2 = F.normalize(a2 - torch.einsum("bi,bi->b", b1, a2).unsqueeze(-1) * b1)
    b3 = torch.cross(b1, b2, dim=1)
    return torch.stack((b1, b2, b3), dim=-1)


def rotmat_to_quat(R: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """rotation_matrix_to_quaternion on [n, 3, 3] (the appended homogeneous column is never read)."""
    r = R.transpose(1, 2)                                    # rmat_t
    d2 = r[:, 2, 2] < eps
    d01 = r[:, 0, 0] > r[:, 1, 1]
    d0n1 = r[:, 0, 0] < -r[:, 1, 1]
    t0 = 1 + r[:, 0, 0] - r[:, 1, 1] - r[:, 2, 2]
    q0 = torch.stack([r[:, 1, 2] - r[:, 2, 1], t0, r[:, 0, 1] + r[:, 1, 0], r[:, 2, 0] + r[:, 0, 2]], -1)
    t1 = 1 - r[:, 0, 0] + r[:, 1, 1] - r[:, 2, 2]
    q1 = torch.stack([r[:, 2, 0] - r[:, 0, 2], r[:, 0, 1] + r[:, 1, 0], t1, r[:, 1, 2] + r[:, 2, 1]], -1)
    t2 = 1 - r[:, 0, 0] - r[:, 1, 1] + r[:, 2, 2]
    q2 = torch.stack([r[:, 0, 1] - r[:, 1, 0], r[:, 2, 0] + r[:, 0, 2], r[:, 1, 2] + r[:, 2, 1], t2], -1)
    t3 = 1 + r[:, 0, 0] + r[:, 1, 1] + r[:, 2, 2]
    q3 = torch.stack([t3, r[:, 1, 2] - r[:, 2, 1], r[:, 2, 0] - r[:, 0, 2], r[:, 0, 1] - r[:, 1, 0]], -1)
    c0 = (d2 & d01).float()[:, None]
    c1 = (d2 & ~d01).float()[:, None]
    c2 = (~d2 & d0n1).float()[:, None]
    c3 = (~d2 & ~d0n1).float()[:, None]
    q = q0 * c0 + q1 * c1 + q2 * c2 + q3 * c3
    q = q / torch.sqrt(t0[:, None] * c0 + t1[:, None] * c1 + t2[:, None] * c2 + t3[:, None] * c3)
    return q * 0.5


def quat_to_axis(q: torch.Tensor) -> torch.Tensor:
    q1, q2, q3 = q[..., 1], q[..., 2], q[..., 3]
    s2 = q1 * q1 + q2 * q2 + q3 * q3
    s = torch.sqrt(s2)
    c = q[..., 0]
    two_theta = 2.0 * torch.where(c < 0.0, torch.atan2(-s, -c), torch.atan2(s, c))
    k = torch.where(s2 > 0.0, two_theta / s, 2.0 * torch.ones_like(s))
    aa = torch.stack([q1 * k, q2 * k, q3 * k], -1)
    aa[torch.isnan(aa)] = 0.0
    return aa


def rot6d_to_axis(rot6d: torch.Tensor) -> torch.Tensor:
    return quat_to_axis(rotmat_to_quat(rot6d_to_rotmat(rot6d)))


# ------------------------------------------------------------------------------------------------ the two smoothers
def smplh_to_smpl_pose(pose: np.ndarray) -> np.ndarray:
    return np.concatenate([pose[:, :69], pose[:, 111:114]], 1) if pose.shape[-1] == 156 else pose


def smooth_smplt(sd, poses: np.ndarray, betas: np.ndarray, trans: np.ndarray, window: int = 64):
    """SMPLTSmoother.preprocess_input -> SmoothNetSMPL -> post_processing.  Returns (poses [L, 72], betas [L, 10], trans [L, 3])."""
    p72 = smplh_to_smpl_pose(poses)
    pose6d = axis_to_rot6d_np(p72.reshape(-1, 3)).reshape(-1, 144)
    seq = torch.from_numpy(np.concatenate([pose6d, betas, trans], 1))
    clips = seq2batches(seq, window)                                            # [B, W, 157]
    init = clips[:, 0:1, 154:157].clone()
    clips[:, :, 154:157] = clips[:, :, 154:157] - init
    x = clips.permute(0, 2, 1)
    with torch.no_grad():
        out = torch.cat([smoothnet_forward(sd, "pose_net.", x[:, :144]), x[:, 144:154].to(torch.float32),
                         smoothnet_forward(sd, "trans_net.", x[:, 154:])], 1).permute(0, 2, 1).clone()
    out[:, :, 154:157] = out[:, :, 154:157] + init.to(out.dtype)
    den = slide_window_mean(out)
    return rot6d_to_axis(den[:, :144].contiguous()).reshape(-1, 72), den[:, 144:154], den[:, 154:157]


def smooth_objrot(sd, rot: np.ndarray, window: int = 64) -> torch.Tensor:
    """ObjrotSmoother.preprocess_input -> SmoothNet -> post_processing.  rot [L, 3, 3] "real" rotation matrices (the loader has
    already transposed the stored ones); returns obj_angles [L, 3, 3] = transposed smoothed rotations, as the reference saves them."""
    rot6d = torch.from_numpy(rot).float().reshape(-1, 3, 3)[:, :, :2].reshape(-1, 6)
    clips = seq2batches(rot6d, window)
    with torch.no_grad():
        out = smoothnet_forward(sd, "", clips.permute(0, 2, 1)).permute(0, 2, 1)
    return rot6d_to_rotmat(slide_window_mean(out)).transpose(1, 2)
