"""Sequence-global stages between the frame-parallel ones, on the device (scripts/demo.sh:28-33, "step 5/7: run SmoothNet + HVOP-Net").

After SIF-Net has run on every rank's frames, ``parallel.gather_trajectory`` assembles the per-frame neural predictions (SURVEY.md 8(e):
``[T, 13]`` = PCA axes 9 | object centre relative to the body 3 | visibility 1).  The reference then runs two scripts with a joblib pack in
between: ``smoothnet/smooth_objrot.py -neural_pca`` (PCA axes -> rotation -> SmoothNet) and ``interp/test_cinfill_autoreg.py`` (HVOP-Net
in-filling of the occluded frames, conditioned on the smoothed SMPL-T motion).  ``object_rotation_stage`` is the same data flow on tensors
that never leave HBM; its output is what ``recon_fit_trivis_full.py -or smooth-hvopnet`` loads as the initial object rotation.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib
from .geom import init_object_orientation
from .infill import CondMotionInfillAutoreg
from .smooth import ObjrotSmoother

P, S = _lib.ptr, _lib.stream_ptr


def pack_neural(pca_axis: torch.Tensor, centers: torch.Tensor, visibility: torch.Tensor) -> torch.Tensor:
    """[t, 13] block a rank contributes to the gather: pca_axis [t,3,3], object centre relative to the body [t,3], visibility [t] or [t,1]."""
    t = pca_axis.shape[0]
    return torch.cat([pca_axis.reshape(t, 9), centers.reshape(t, 3), visibility.reshape(t, 1)], 1).float().contiguous()


def smpl_rot6d(poses: torch.Tensor, trans: torch.Tensor):
    """``prep_smpl_rot6d`` (interp/test_infiller.py:186-194): SMPL-H poses reduced to the 24 SMPL joints, axis-angle -> 6-D.  Returns
    (rot6d_smpl [T,144], trans [T,3]) as device tensors."""
    T = poses.shape[0]
    poses, trans = poses.float().contiguous(), trans.float().contiguous()
    seq = torch.empty(T, 157, device=poses.device)
    zeros = torch.zeros(T, 10, device=poses.device)
    with torch.cuda.device(poses.device):
        _lib.call("vt_smooth_pack_smplt", P(poses), poses.shape[1], P(zeros), P(trans), T, P(seq), S())
    return seq[:, :144].contiguous(), seq[:, 154:157].contiguous()


def object_rotation_stage(neural: torch.Tensor, smpl_poses: torch.Tensor, smpl_trans: torch.Tensor, obj_trans: torch.Tensor, template_axes: torch.Tensor,
                          smoother: ObjrotSmoother, infiller: CondMotionInfillAutoreg, occ_thres: float = 0.5) -> Dict[str, Optional[torch.Tensor]]:
    """neural [T,13] (``pack_neural`` rows, all-gathered), smpl_poses [T, 72|156] + smpl_trans [T,3] (the smoothed SMPL-T fit), obj_trans
    [T,3], template_axes [3,3] (``PCAUtil.compute_pca`` of the object template).

    1. ``PCAUtil.init_object_orientation`` (no noise) and the transpose of smooth_objrot.py:46-58 -> 'obj_rot'
    2. ``ObjrotSmoother``: SmoothNet over windows of 64 -> 'obj_angles' (R^T, as stored)
    3. ``numpy_rotmat_to_6d(obj_angles^T)`` (interp/test_infiller.py:183) and ``prep_smpl_rot6d``
    4. ``CondMotionInfillAutoreg``: visibility = neural[:, 12] < occ_thres is in-filled autoregressively
    Returns {'obj_angles_smooth', 'obj_angles', 'obj_trans', 'obj_scales', 'infilled'}; when HVOP-Net skips the sequence (no visible seed
    frames) 'obj_angles' is the smoothed input, as the reference re-saves it."""
    T = neural.shape[0]
    R0 = init_object_orientation(neural[:, :9].reshape(T, 3, 3), template_axes, no_rand=True)
    angles_s = smoother.smooth(R0.transpose(1, 2).contiguous())
    rot6d_obj = angles_s.transpose(1, 2)[:, :, :2].reshape(T, 6).contiguous()
    rot6d_smpl, trans_smpl = smpl_rot6d(smpl_poses, smpl_trans)
    res = infiller.infill(rot6d_smpl, trans_smpl, rot6d_obj, obj_trans, neural[:, 12], occ_thres=occ_thres)
    if res is None:
        return {"obj_angles_smooth": angles_s, "obj_angles": angles_s, "obj_trans": obj_trans.clone(), "obj_scales": torch.ones(T, device=neural.device),
                "infilled": False}
    return {"obj_angles_smooth": angles_s, "obj_angles": res["obj_angles"], "obj_trans": res["obj_trans"], "obj_scales": res["obj_scales"], "infilled": True}
