"""CPU restatement of the SMPL-H layer (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Follows lib_smpl/smplpytorch/smplpytorch/pytorch/smpl_layer.py:73-176 (forward), rodrigues_layer.py:13-52 (axis-angle ->
quaternion -> rotation, with eps added to the VECTOR before the norm) and tensutils.py:6-53.  Vectorised over joints
instead of the reference's per-joint Python loops; differentiable through torch autograd, fp32 or fp64.
"""
from __future__ import annotations

import torch


def rodrigues(axisang: torch.Tensor) -> torch.Tensor:
    """[..., 3] -> [..., 3, 3]; batch_rodrigues + quat2mat (rodrigues_layer.py:13-52)."""
    angle = torch.norm(axisang + 1e-8, p=2, dim=-1, keepdim=True)
    axis = axisang / angle
    half = angle * 0.5
    quat = torch.cat([torch.cos(half), torch.sin(half) * axis], -1)
    quat = quat / quat.norm(p=2, dim=-1, keepdim=True)
    w, x, y, z = quat.unbind(-1)
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], -1).reshape(*axisang.shape[:-1], 3, 3)


def smpl_forward(model, pose, betas, trans, offsets=None, scale: float = 1.0):
    """SMPL_Layer.forward.  model: dict of th_* buffers + 'parents'.  Returns (verts [B,V,3], jtr [B,J,3], v_posed, naked)."""
    B = pose.shape[0]
    dt = pose.dtype
    parents = model["parents"]
    J = len(parents)
    R = rodrigues(pose.reshape(B, J, 3))                                            # th_posemap_axisang
    pose_map = (R[:, 1:] - torch.eye(3, dtype=dt)).reshape(B, (J - 1) * 9)            # subtract_flat_id
    v_shaped = model["th_v_template"].to(dt) + torch.matmul(model["th_shapedirs"].to(dt), betas.t()).permute(2, 0, 1)
    joints = torch.matmul(model["th_J_regressor"].to(dt), v_shaped)                   # :95
    naked = v_shaped + torch.matmul(model["th_posedirs"].to(dt), pose_map.t()).permute(2, 0, 1)
    v_posed = naked + offsets if offsets is not None else naked
    bottom = torch.tensor([0.0, 0.0, 0.0, 1.0], dtype=dt).expand(B, 1, 4)
    G = [torch.cat([torch.cat([R[:, 0], joints[:, 0, :, None]], 2), bottom], 1)]      # :111-113
    for i in range(1, J):
        rel = torch.cat([torch.cat([R[:, i], (joints[:, i] - joints[:, parents[i]])[:, :, None]], 2), bottom], 1)
        G.append(torch.matmul(G[parents[i]], rel))                                   # :116-123
    G = torch.stack(G, 1)                                                             # [B,J,4,4]
    jh = torch.cat([joints, torch.zeros(B, J, 1, dtype=dt)], 2)
    A = G.clone()
    A[:, :, :, 3] = G[:, :, :, 3] - torch.einsum("bjrc,bjc->bjr", G, jh)             # th_results - th_pack(G @ [j;0]) :129-137
    T = torch.einsum("bjrc,vj->bvrc", A, model["th_weights"].to(dt))                  # :139
    vh = torch.cat([v_posed, torch.ones(B, v_posed.shape[1], 1, dtype=dt)], 2)
    verts = torch.einsum("bvrc,bvc->bvr", T, vh)[:, :, :3] * scale
    jtr = G[:, :, :3, 3] * scale
    return verts + trans[:, None], jtr + trans[:, None], v_posed, naked


def landmarks(reg_indices, reg_values, reg_shape, verts):
    """batch_sparse_dense_matmul (lib_smpl/torch_functions.py:52-76) for one COO regressor [V, L]: J = reg^T verts."""
    dense = torch.zeros(reg_shape, dtype=verts.dtype)
    dense.index_put_((reg_indices[0], reg_indices[1]), reg_values.to(verts.dtype), accumulate=True)
    return torch.einsum("vl,bvc->blc", dense, verts)
