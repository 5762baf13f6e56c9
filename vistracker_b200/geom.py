"""Differentiable wrappers of the joint-optimisation helpers (csrc/geom.cu): SO(3) projection, rigid object transform and the
ragged Chamfer distance of the contact loss."""
from __future__ import annotations

from typing import List

import torch

from . import _lib

P, S = _lib.ptr, _lib.stream_ptr


class _So3Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mat):
        m = mat.detach().float().contiguous()
        out = torch.empty_like(m)
        with torch.cuda.device(m.device):
            _lib.call("vt_so3_project_fwd", P(m), m.shape[0], P(out), S())
        ctx.save_for_backward(m)
        return out

    @staticmethod
    def backward(ctx, g):
        (m,) = ctx.saved_tensors
        g = g.float().contiguous()
        gm = torch.empty_like(m)
        with torch.cuda.device(m.device):
            _lib.call("vt_so3_project_bwd", P(m), P(g), m.shape[0], P(gm), S())
        return gm


def project_so3(mat: torch.Tensor) -> torch.Tensor:
    """ReconFitterBase.project_so3 (recon/recon_fit_base.py:178-199) for mat [B, 3, 3] on a CUDA device."""
    assert mat.shape[1:] == (3, 3), f"invalid shape {mat.shape}"
    if not mat.is_cuda:
        raise RuntimeError("vistracker_b200 has no CPU path")
    return _So3Fn.apply(mat)


def decopose_axis(rot: torch.Tensor, no_rand: bool = False, noise: torch.Tensor = None) -> torch.Tensor:
    """ReconFitterBase.decopose_axis (recon_fit_base.py:461-469): SO(3) projection of ``rot + 1e-4 * U(0,1)``.  ``noise``
    lets the caller inject the uniform tensor (parity runs replay the oracle's draws)."""
    if no_rand:
        return project_so3(rot)
    if noise is None:
        noise = torch.rand(rot.shape[0], 3, 3, device=rot.device)
    return project_so3(rot + 1e-4 * noise)


def init_object_orientation(tgt_axis: torch.Tensor, src_axis: torch.Tensor, no_rand: bool = False, noise: torch.Tensor = None) -> torch.Tensor:
    """``ReconFitterBase.init_object_orientation`` (recon/recon_fit_base.py:202-216): the rotation from the template's PCA axes ``src_axis``
    ([B,3,3] or one [3,3]) to the predicted ones ``tgt_axis`` [B,3,3]: ``decopose_axis(pinv(src) @ tgt)``.  ``no_rand=True`` is
    ``PCAUtil.init_object_orientation`` (recon/pca_util.py:59-72, used by the object smoother and HVOP-Net inputs): no noise."""
    t = tgt_axis.detach().float().contiguous()
    s = src_axis.detach().to(t.device).float().contiguous()
    B = t.shape[0]
    if t.shape[1:] != (3, 3) or s.shape[-2:] != (3, 3) or (s.dim() == 3 and s.shape[0] != B):
        raise ValueError(f"invalid shapes {tuple(tgt_axis.shape)} / {tuple(src_axis.shape)}")
    if not no_rand and noise is None:
        noise = torch.rand(B, 3, 3, device=t.device)
    nz = None if no_rand else noise.to(t.device).float().contiguous()
    out = torch.empty_like(t)
    with torch.cuda.device(t.device):
        _lib.call("vt_pca_orientation", P(t), P(s), int(s.dim() == 3), P(nz), B, P(out), S())
    return out


class _ChamferFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, x_off, y_off):
        xc, yc = x.detach().float().contiguous(), y.detach().float().contiguous()
        n = x_off.numel() - 1
        nn_x = torch.empty(xc.shape[0], dtype=torch.int32, device=xc.device)
        nn_y = torch.empty(yc.shape[0], dtype=torch.int32, device=xc.device)
        loss = torch.empty(1, dtype=torch.float32, device=xc.device)
        with torch.cuda.device(xc.device):
            _lib.call("vt_chamfer_fwd", P(xc), P(x_off), P(yc), P(y_off), n, P(nn_x), P(nn_y), P(loss), S())
        ctx.save_for_backward(xc, yc, x_off, y_off, nn_x, nn_y)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        xc, yc, x_off, y_off, nn_x, nn_y = ctx.saved_tensors
        gx, gy = torch.zeros_like(xc), torch.zeros_like(yc)
        gl = g.reshape(1).float().contiguous()
        with torch.cuda.device(xc.device):
            _lib.call("vt_chamfer_bwd", P(xc), P(x_off), P(yc), P(y_off), x_off.numel() - 1, P(nn_x), P(nn_y), P(gl), P(gx), P(gy), S())
        return gx, gy, None, None


def chamfer_distance_ragged(xs: List[torch.Tensor], ys: List[torch.Tensor]) -> torch.Tensor:
    """pytorch3d.loss.chamfer_distance(Pointclouds(xs), Pointclouds(ys))[0] with default reductions, for lists of [n_i, 3]
    clouds (recon/recon_fit_trivis_full.py:452-456).  Differentiable w.r.t. every cloud."""
    assert len(xs) == len(ys) and len(xs) > 0
    dev = xs[0].device
    off = lambda cl: torch.tensor([0] + list(torch.tensor([c.shape[0] for c in cl]).cumsum(0)), dtype=torch.int32, device=dev)
    return _ChamferFn.apply(torch.cat(xs, 0), torch.cat(ys, 0), off(xs), off(ys))


def chamfer_distance_packed(x: torch.Tensor, y: torch.Tensor, x_off: torch.Tensor, y_off: torch.Tensor) -> torch.Tensor:
    """Same operator on already-packed clouds: x [sum n_i, 3], y [sum m_i, 3] with int32 row offsets [N + 1] each."""
    return _ChamferFn.apply(x, y, x_off, y_off)


def eval_chamfer_distance(x: torch.Tensor, y: torch.Tensor, direction: str = "bi") -> torch.Tensor:
    """``chamfer_distance`` of recon/eval/chamfer_distance.py:10-52 for batches of dense clouds x [B, n, 3], y [B, m, 3] (or single
    [n, 3] clouds): mean Euclidean nearest-neighbour distance, 'x_to_y', 'y_to_x' or 'bi' (the sum of both) -- one value per frame,
    in the input's units (the reference reports it x100 as centimetres, recon/eval/evaluate.py:126-154)."""
    from . import _lib
    if direction not in ("bi", "x_to_y", "y_to_x"):
        raise ValueError("Invalid direction type. Supported types: 'y_x', 'x_y', 'bi'")
    single = x.dim() == 2
    xb, yb = (x[None] if single else x).float().contiguous(), (y[None] if single else y).float().contiguous()
    if not xb.is_cuda:
        raise RuntimeError("vistracker_b200 has no CPU path: eval_chamfer_distance needs CUDA tensors")
    B, n, m = xb.shape[0], xb.shape[1], yb.shape[1]
    dx, dy = torch.empty(B, n, device=xb.device), torch.empty(B, m, device=xb.device)
    with torch.cuda.device(xb.device):
        _lib.call("vt_nn_dist", _lib.ptr(xb), n, _lib.ptr(yb), m, B, _lib.ptr(dx), _lib.ptr(dy), _lib.stream_ptr())
    out = {"x_to_y": dx.mean(1), "y_to_x": dy.mean(1)}
    res = out["x_to_y"] + out["y_to_x"] if direction == "bi" else out[direction]
    return res[0] if single else res


def procrustes_transform(S1: torch.Tensor, S2: torch.Tensor):
    """``compute_transform`` (recon/eval/pose_utils.py:153-198) for clouds S1, S2 [N, 3] or batches [B, N, 3]: returns (R [.., 3, 3],
    t [.., 3], scale [..]) with ``scale * R @ p + t`` mapping S1 onto S2 -- what ``ProcrusteAlign.get_transform`` computes from the
    combined SMPL + object vertices of an alignment window (pose_utils.py:49-69, evalvideo_packed.py:108-126)."""
    from . import _lib
    single = S1.dim() == 2
    a, b = (S1[None] if single else S1).float().contiguous(), (S2[None] if single else S2).float().contiguous()
    if not a.is_cuda:
        raise RuntimeError("vistracker_b200 has no CPU path: procrustes_transform needs CUDA tensors")
    if a.shape != b.shape:
        raise ValueError(f"point sets must correspond: {tuple(S1.shape)} vs {tuple(S2.shape)}")
    B, N = a.shape[0], a.shape[1]
    ws = torch.empty(B, 16, dtype=torch.float64, device=a.device)
    R, t, s = torch.empty(B, 3, 3, device=a.device), torch.empty(B, 3, device=a.device), torch.empty(B, device=a.device)
    with torch.cuda.device(a.device):
        _lib.call("vt_procrustes", _lib.ptr(a), _lib.ptr(b), N, B, _lib.ptr(ws), _lib.ptr(R), _lib.ptr(t), _lib.ptr(s), _lib.stream_ptr())
    return (R[0], t[0], s[0]) if single else (R, t, s)


def apply_similarity(points: torch.Tensor, R: torch.Tensor, t: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """``scale * R @ p + t`` for points [n, 3] (single transform) or [B, n, 3] (one transform per cloud)."""
    from . import _lib
    single = points.dim() == 2
    p = (points[None] if single else points).float().contiguous()
    B, n = p.shape[0], p.shape[1]
    R, t, scale = R.reshape(B, 9).float().contiguous(), t.reshape(B, 3).float().contiguous(), scale.reshape(B).float().contiguous()
    out = torch.empty_like(p)
    with torch.cuda.device(p.device):
        _lib.call("vt_similarity_apply", _lib.ptr(p), n, B, _lib.ptr(R), _lib.ptr(t), _lib.ptr(scale), _lib.ptr(out), _lib.stream_ptr())
    return out[0] if single else out


def compute_pca(points, sign: str = "u"):
    """``PCAUtil.compute_pca`` / ``compute_pca_init`` (recon/pca_util.py:12-24, recon/recon_fit_base.py:130-144): the three principal axes of
    the object template's vertices as rows (``sklearn.decomposition.PCA(3).fit(points).components_``) -- the ``src_axis`` of
    ``init_object_orientation``.  Host arithmetic, once per object.  scikit-learn fixes the sign of every axis with ``svd_flip``; its rule
    changed in scikit-learn 1.5: ``sign='u'`` (before: the sample with the largest absolute projection on the axis projects positively --
    what the released checkpoints were trained against, requirements.txt does not pin the version) or ``sign='v'`` (since: the largest
    absolute coordinate of the axis is positive)."""
    import numpy as np
    X = np.asarray(points.detach().cpu() if torch.is_tensor(points) else points, dtype=np.float64)
    Xc = X - X.mean(0)
    U, S, Vt = np.linalg.svd(Xc, full_matrices=False)
    if sign == "u":
        s = np.sign(U[np.argmax(np.abs(U), axis=0), np.arange(U.shape[1])])
    elif sign == "v":
        s = np.sign(Vt[np.arange(Vt.shape[0]), np.argmax(np.abs(Vt), axis=1)])
    else:
        raise ValueError("sign must be 'u' or 'v'")
    s[s == 0] = 1
    return Vt * s[:, None]
