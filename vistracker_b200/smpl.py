"""Drop-in for ``SMPL_Layer`` (lib_smpl/smplpytorch/smplpytorch/pytorch/smpl_layer.py:20-208) and the landmark regressors
(lib_smpl/wrapper_pytorch.py:187-203, lib_smpl/torch_functions.py:52-76, lib_smpl/body_landmark.py:16-28) on B200.

``SMPL_Layer.forward(th_pose_axisang, th_betas, th_trans, th_offsets=None, scale=1.)`` returns
``(verts, jtr, v_posed, naked)`` and is differentiable w.r.t. pose, betas and trans (and offsets); forward is 3 kernel
launches, backward 3 (csrc/smpl.cu) instead of ~1.5k torch ops each way.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib


class _SmplFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, layer: "SMPL_Layer", pose, betas, trans, offsets, scale):
        B, dev, m = pose.shape[0], layer.device, layer
        f = dict(dtype=torch.float32, device=dev)
        pose_c, betas_c, trans_c = (t.detach().to(**f).contiguous() for t in (pose, betas, trans))
        off_c = None if offsets is None else offsets.detach().to(**f).contiguous()
        coef = torch.empty(B, m.kdp, **f); R = torch.empty(B, m.J, 9, **f); J = torch.empty(B, m.J, 3, **f)
        G = torch.empty(B, m.J, 12, **f); A = torch.empty(B, m.J, 12, **f)
        naked = torch.empty(B, m.V, 3, **f); verts = torch.empty(B, m.V, 3, **f); jtr = torch.empty(B, m.J, 3, **f)
        v_posed = torch.empty(B, m.V, 3, **f) if off_c is not None else naked
        with torch.cuda.device(dev):
            _lib.call("vt_smpl_fwd", ctypes.byref(m.struct), _lib.ptr(pose_c), _lib.ptr(betas_c), _lib.ptr(trans_c), _lib.ptr(off_c),
                      float(scale), B, _lib.ptr(coef), _lib.ptr(R), _lib.ptr(J), _lib.ptr(G), _lib.ptr(A), _lib.ptr(naked),
                      _lib.ptr(v_posed), _lib.ptr(verts), _lib.ptr(jtr), _lib.stream_ptr())
        ctx.layer, ctx.scale, ctx.has_off = layer, float(scale), off_c is not None
        ctx.save_for_backward(pose_c, R, J, G, A, v_posed)
        if off_c is None:
            ctx.mark_non_differentiable(naked)
            return verts, jtr, None, naked
        ctx.mark_non_differentiable(v_posed, naked)
        return verts, jtr, v_posed, naked

    @staticmethod
    def backward(ctx, g_verts, g_jtr, _g_vp, _g_naked):
        pose_c, R, J, G, A, v_posed = ctx.saved_tensors
        m, B = ctx.layer, pose_c.shape[0]
        f = dict(dtype=torch.float32, device=m.device)
        gv = None if g_verts is None else g_verts.to(**f).contiguous()
        gj = None if g_jtr is None else g_jtr.to(**f).contiguous()
        g_vposed = torch.empty(B, m.nv3p, **f); gA = torch.empty(B, m.J, 12, **f); g_coef = torch.empty(B, m.kdp, **f)
        g_ts = torch.empty(B, 3, **f)
        g_pose = torch.empty(B, m.J * 3, **f); g_betas = torch.empty(B, m.n_betas, **f); g_trans = torch.empty(B, 3, **f)
        with torch.cuda.device(m.device):
            _lib.call("vt_smpl_bwd", ctypes.byref(m.struct), _lib.ptr(pose_c), _lib.ptr(R), _lib.ptr(J), _lib.ptr(G), _lib.ptr(A),
                      _lib.ptr(v_posed), _lib.ptr(gv), _lib.ptr(gj), ctx.scale, B, _lib.ptr(g_vposed), _lib.ptr(gA), _lib.ptr(g_coef),
                      _lib.ptr(g_ts), _lib.ptr(g_pose), _lib.ptr(g_betas), _lib.ptr(g_trans), _lib.stream_ptr())
        g_off = g_vposed[:, :3 * m.V].reshape(B, m.V, 3) if (ctx.has_off and gv is not None) else None
        return None, g_pose, g_betas, g_trans, g_off, None


class SMPL_Layer:
    """B200 SMPL(-H) layer.  Construct from the reference's buffers:

        SMPL_Layer.from_buffers({'th_v_template', 'th_shapedirs', 'th_posedirs', 'th_J_regressor', 'th_weights',
                                 'th_betas', 'th_faces'}, kintree_parents, device)

    (the reference constructor reads the licensed SMPLH_{gender}.pkl through chumpy; any loader that yields those arrays
    works, e.g. ``reference_layer.state_dict()`` + ``reference_layer.kintree_parents``)."""

    def __init__(self, center_idx=None, gender="neutral", model_root="smpl/native/models", num_betas=300, hands=False):
        raise NotImplementedError("use SMPL_Layer.from_buffers(...): SMPLH_*.pkl is not redistributable and needs chumpy to read")

    @classmethod
    def from_buffers(cls, buffers: Dict[str, torch.Tensor], kintree_parents, device="cuda:0", hands=True, center_idx=None,
                     gender="male") -> "SMPL_Layer":
        self = object.__new__(cls)
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("vistracker_b200 has no CPU path")
        _lib.load()
        self.device, self.hands, self.center_idx, self.gender = dev, hands, center_idx, gender
        self.kintree_parents = list(kintree_parents)
        self.num_joints = len(self.kintree_parents)
        # the model buffers as the reference holds them (float64 pickles) -> kernel layouts: host C behind the ABI (vt_pack_weights_smpl)
        h64 = lambda t: t.detach().to("cpu", torch.float64).contiguous()
        vt, sd, pd, jr = h64(buffers["th_v_template"]).reshape(-1, 3), h64(buffers["th_shapedirs"]), h64(buffers["th_posedirs"]), h64(buffers["th_J_regressor"])
        w = buffers["th_weights"].detach().to("cpu", torch.float32).contiguous()
        V, J, nb = vt.shape[0], self.num_joints, sd.shape[2]
        assert pd.shape == (V, 3, 9 * (J - 1)) and jr.shape == (J, V) and w.shape == (V, J)
        hp = lambda t: ctypes.c_void_p(t.data_ptr())
        dims = [ctypes.c_int() for _ in range(4)]
        _lib.call("vt_smpl_pack_dims", V, J, nb, hp(w), *(ctypes.byref(d) for d in dims))
        self.V, self.J, self.n_betas = V, J, nb
        self.kd, self.kdp, self.nv3p, nnz = (d.value for d in dims)
        parents = torch.tensor(self.kintree_parents, dtype=torch.int32)
        e = lambda *sh, dt=torch.float32: torch.empty(*sh, dtype=dt)
        host = dict(templ=e(3 * V), dirs=e(self.kdp, self.nv3p), dirsT=e(self.nv3p, self.kdp), j_templ=e(J, 3), j_dirs=e(J, 3, nb),
                    parents=e(J, dt=torch.int32), skin_idx=e(V, nnz, dt=torch.int32), skin_w=e(V, nnz))
        _lib.call("vt_pack_weights_smpl", hp(vt), hp(sd), hp(pd), hp(jr), hp(w), hp(parents), V, J, nb,
                  *(hp(host[k]) for k in ("templ", "dirs", "dirsT", "j_templ", "j_dirs", "parents", "skin_idx", "skin_w")))
        self._keep = {k: t.to(dev) for k, t in host.items()}
        self.struct = _lib.SmplModelStruct(V, J, nb, self.kd, self.kdp, self.nv3p, nnz,
                                           *(ctypes.c_void_p(self._keep[k].data_ptr()) for k in
                                             ("templ", "dirs", "dirsT", "j_templ", "j_dirs", "parents", "skin_idx", "skin_w")))
        # reference buffer names, for callers that poke at them
        self.th_betas = buffers.get("th_betas", torch.zeros(1, nb)).to(dev)
        self.th_faces = buffers["th_faces"].to(dev) if "th_faces" in buffers else None
        self.th_v_template, self.th_weights = buffers["th_v_template"].to(dev), buffers["th_weights"].to(dev)
        self.faces = None if self.th_faces is None else self.th_faces.cpu().numpy().astype(np.int32)
        return self

    def to(self, device):
        if torch.device(device) != self.device:
            raise RuntimeError("rebuild the layer with from_buffers(..., device=...) to move it")
        return self

    def __call__(self, *a, **k):
        return self.forward(*a, **k)

    def forward(self, th_pose_axisang, th_betas=None, th_trans=None, th_offsets=None, scale=1.):
        """(verts [B,V,3], jtr [B,J,3], v_posed [B,V,3], naked [B,V,3]) -- smpl_layer.py:73-176."""
        B = th_pose_axisang.shape[0]
        if th_betas is None:
            th_betas = self.th_betas.expand(B, -1)
        if th_trans is None:
            th_trans = torch.zeros(B, 3, device=self.device)
        if th_pose_axisang.shape[1] != 3 * self.J or th_betas.shape[-1] != self.n_betas:
            raise ValueError(f"expected pose [B,{3 * self.J}] and betas [B,{self.n_betas}]")
        verts, jtr, v_posed, naked = _SmplFn.apply(self, th_pose_axisang, th_betas, th_trans, th_offsets, scale)
        return verts, jtr, (naked if v_posed is None else v_posed), naked

    def get_root_joint(self, th_pose_axisang, th_betas, th_trans):
        """smpl_layer.py:178-208: root joint = J_0(betas) + trans, [B, 1, 3]."""
        j0 = self._keep["j_templ"][0] + torch.einsum("ck,bk->bc", self._keep["j_dirs"][0], th_betas.to(self.device).float())
        return (j0 + th_trans.to(self.device))[:, None, :]


class LandmarkRegressor:
    """A sparse [V x L] regressor (assets/{body25,face,hand}_regressor.pkl: COO indices [2,nnz], values, shape) applied
    as J = reg^T verts for a whole batch in one launch (the reference loops ``torch.sparse.mm`` over the batch)."""

    def __init__(self, indices, values, shape, device="cuda:0"):
        idx = torch.as_tensor(np.asarray(indices)).long()
        val = torch.as_tensor(np.asarray(values)).float()
        V, L = int(shape[0]), int(shape[1])
        order = torch.argsort(idx[1] * V + idx[0])
        rows, cols, val = idx[1][order], idx[0][order], val[order]
        rowptr = torch.zeros(L + 1, dtype=torch.int64)
        rowptr[1:] = torch.cumsum(torch.bincount(rows, minlength=L), 0)
        self.V, self.L, self.device = V, L, torch.device(device)
        self.rowptr, self.col, self.val = (rowptr.to(torch.int32).to(self.device), cols.to(torch.int32).to(self.device),
                                           val.to(self.device))

    def __call__(self, verts):
        return _LandmarkFn.apply(self, verts)


class _LandmarkFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, reg: LandmarkRegressor, verts):
        v = verts.detach().to(reg.device, torch.float32).contiguous()
        B = v.shape[0]
        out = torch.empty(B, reg.L, 3, dtype=torch.float32, device=reg.device)
        with torch.cuda.device(reg.device):
            _lib.call("vt_landmarks_fwd", _lib.ptr(v), B, reg.V, _lib.ptr(reg.rowptr), _lib.ptr(reg.col), _lib.ptr(reg.val), reg.L,
                      _lib.ptr(out), _lib.stream_ptr())
        ctx.reg, ctx.B = reg, B
        return out

    @staticmethod
    def backward(ctx, g):
        reg = ctx.reg
        g = g.to(reg.device, torch.float32).contiguous()
        gv = torch.zeros(ctx.B, reg.V, 3, dtype=torch.float32, device=reg.device)
        with torch.cuda.device(reg.device):
            _lib.call("vt_landmarks_bwd", _lib.ptr(g), ctx.B, reg.V, _lib.ptr(reg.rowptr), _lib.ptr(reg.col), _lib.ptr(reg.val), reg.L,
                      _lib.ptr(gv), _lib.stream_ptr())
        return None, gv
