"""Rasteriser front-ends with the reference's object surface (csrc/raster.cu):

* ``SilhouetteRenderer`` -- what ``nr.renderer.Renderer(image_size, K=K_roi, R=I, t=0, orig_size=1, anti_aliasing=False)(verts,
  faces, mode='silhouettes')`` computes in ``SilLossROI`` (recon/obj_pose_roi.py:77-94,183-202), differentiable w.r.t. the vertices;
* ``SilLossROI.forward`` -- the occlusion-aware mask loss on pre-cropped ROI masks;
* ``TriplaneNrRenderer.render_3views`` -- the three orthographic occupancy masks of render/render_triplane_nr.py:88-139.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

P, S = _lib.ptr, _lib.stream_ptr


def _cull_ws(B: int, F: int, device):
    """Workspace for the per-face / per-chunk bounding boxes of vt_raster_fwd (``VT_RASTER_CULL=0``: every tile scans every face)."""
    import os
    if os.environ.get("VT_RASTER_CULL", "1") == "0":
        return None
    return torch.empty(_lib.load().vt_raster_cull_floats(B, F), device=device, dtype=torch.float32)


class _SilFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, r: "SilhouetteRenderer", verts):
        v = verts.detach().float().contiguous()
        B, V = v.shape[0], v.shape[1]
        F, isz, dev = r.faces.shape[0], r.image_size, v.device
        faces_ndc = torch.empty(B, 2 * F, 9, device=dev)
        fidx = torch.empty(B, isz, isz, dtype=torch.int32, device=dev)
        alpha = torch.empty(B, isz, isz, device=dev)
        cull = _cull_ws(B, F, dev)
        with torch.cuda.device(dev):
            _lib.call("vt_raster_fwd", P(v), P(r.faces), B, V, F, r.mode, P(r.K4), isz, P(faces_ndc), P(fidx), P(alpha), None, P(cull), S())
        ctx.r = r
        ctx.save_for_backward(v, faces_ndc, fidx, alpha)
        return alpha

    @staticmethod
    def backward(ctx, g):
        v, faces_ndc, fidx, alpha = ctx.saved_tensors
        r = ctx.r
        B, V, F = v.shape[0], v.shape[1], r.faces.shape[0]
        g = g.float().contiguous()
        g_faces = torch.empty(B, 2 * F, 9, device=v.device)
        g_verts = torch.empty_like(v)
        with torch.cuda.device(v.device):
            _lib.call("vt_raster_bwd", P(v), P(r.faces), B, V, F, r.mode, P(r.K4), r.image_size, P(faces_ndc), P(fidx), P(alpha), P(g),
                      P(g_faces), P(g_verts), S())
        return None, g_verts


class SilhouetteRenderer:
    """faces [F, 3] (shared by the batch); K [B, 3, 3] normalised ROI intrinsics (projection mode) or None (orthographic)."""

    def __init__(self, faces, image_size=256, K=None, device="cuda:0"):
        dev = torch.device(device)
        self.faces = torch.as_tensor(np.asarray(faces)).to(torch.int32).contiguous().to(dev)
        self.image_size, self.mode = int(image_size), 0 if K is not None else 1
        self.K4 = None
        if K is not None:
            K = torch.as_tensor(K).float().to(dev)
            self.K4 = torch.stack([K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2]], 1).contiguous()

    def __call__(self, verts):
        return _SilFn.apply(self, verts)

    def render_depth(self, verts):
        """[B, S, S] depth (far = 100 where nothing is hit); not differentiable (used for occupancy only)."""
        v = verts.detach().float().contiguous()
        B, V, F = v.shape[0], v.shape[1], self.faces.shape[0]
        faces_ndc = torch.empty(B, 2 * F, 9, device=v.device)
        fidx = torch.empty(B, self.image_size, self.image_size, dtype=torch.int32, device=v.device)
        depth = torch.empty(B, self.image_size, self.image_size, device=v.device)
        cull = _cull_ws(B, F, v.device)
        with torch.cuda.device(v.device):
            _lib.call("vt_raster_fwd", P(v), P(self.faces), B, V, F, self.mode, P(self.K4), self.image_size, P(faces_ndc), P(fidx), None,
                      P(depth), P(cull), S())
        return depth


class SilLossROI:
    """Occlusion-aware silhouette loss of recon/obj_pose_roi.py on already-cropped ROI masks.

    keep_mask / image_ref: [B, S, S] as ``cvt_masks`` / ``(obj > 0)`` produce them (obj_pose_roi.py:62-75,157-170);
    K_roi: [B, 3, 3] from ``compute_K_roi`` (:123-155); vertices [V, 3] / faces [F, 3] of the centred template."""

    def __init__(self, keep_mask, image_ref, K_roi, vertices, faces, rend_size=256, device="cuda:0"):
        dev = torch.device(device)
        self.keep_mask, self.image_ref = keep_mask.float().to(dev), image_ref.float().to(dev)
        self.vertices = torch.as_tensor(np.asarray(vertices), dtype=torch.float32).to(dev)
        self.renderer = SilhouetteRenderer(faces, rend_size, K_roi, dev)

    @staticmethod
    def compute_K_roi(bbox_square, image_width=2048, fx=979.7844, fy=979.840, cx=1018.952, cy=779.486):
        """obj_pose_roi.py:123-155: intrinsics of the square ROI (x, y, b, b) in the original image, normalised to the ROI."""
        x, y, b, w = bbox_square
        assert b == w, "the given bbox is not square!"
        if fx > 1.0:
            fx, fy, cx, cy = fx / image_width, fy / image_width, cx / image_width, cy / image_width
        return torch.tensor([[fx * image_width / b, 0, (cx * image_width - x) / b], [0, fy * image_width / b, (cy * image_width - y) / b],
                             [0, 0, 1]], dtype=torch.float32)

    def apply_transformation(self, R, obj_t, obj_s):
        verts = torch.matmul(self.vertices.unsqueeze(0), R) + obj_t.unsqueeze(1)
        return obj_s.view(-1, 1, 1) * verts

    def forward(self, R, obj_t, obj_s, reduction="mean"):
        verts = self.apply_transformation(R, obj_t, obj_s)
        image = self.keep_mask * self.renderer(verts)
        per_frame = torch.sum((image - self.image_ref) ** 2, dim=(1, 2))
        if reduction == "mean":
            return {"mask": per_frame.mean()}, image
        if reduction == "none":
            return {"mask": per_frame}, image
        raise NotImplementedError(f"Unknown reduction type: {reduction}")

    __call__ = forward


class TriplaneNrRenderer:
    """render/render_triplane_nr.py:25-30,88-139.  nr.Renderer's default anti_aliasing=True renders at 2x and average-pools the
    depth, so ``depth < far`` marks a pixel as soon as ANY of its 2x2 sub-samples is covered."""

    def __init__(self, image_size=512, device="cuda:0"):
        self.image_size, self.device, self.z_offset = image_size, torch.device(device), 10.0

    @staticmethod
    def transform_view(points_center, view, z_offset=10.0):
        x, y, z = points_center[..., 0], points_center[..., 1], points_center[..., 2]
        if view == "right":
            return torch.stack([z, -y, -x + z_offset], -1)
        if view == "back":
            return torch.stack([-x, -y, -z + z_offset], -1)
        if view == "top":
            return torch.stack([x, z, y + z_offset], -1)
        raise AssertionError(view)

    def render_3views(self, faces, points_center):
        """points_center [B, V, 3] (centred on body-25 joint 8) -> uint8 masks [B, 3, S, S] ordered right, back, top."""
        pts = torch.as_tensor(points_center, dtype=torch.float32).to(self.device)
        if pts.dim() == 2:
            pts = pts[None]
        r = SilhouetteRenderer(faces, 2 * self.image_size, None, self.device)
        out = []
        for view in ("right", "back", "top"):
            depth = r.render_depth(self.transform_view(pts, view, self.z_offset))
            hit = (depth < 100.0).float()
            out.append(torch.nn.functional.max_pool2d(hit[:, None], 2)[:, 0] > 0)
        return torch.stack(out, 1).to(torch.uint8)
