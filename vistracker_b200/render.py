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
            ws = torch.empty(_lib.load().vt_workspace_bytes_raster_bwd(B, r.image_size), dtype=torch.uint8, device=v.device) if r.skip_walks else None
            _lib.call("vt_raster_bwd_ws", P(v), P(r.faces), B, V, F, r.mode, P(r.K4), r.image_size, P(faces_ndc), P(fidx), P(alpha), P(g),
                      P(g_faces), P(g_verts), P(ws), S())
        return None, g_verts


class SilhouetteRenderer:
    """faces [F, 3] (shared by the batch); K [B, 3, 3] normalised ROI intrinsics (projection mode) or None (orthographic)."""

    def __init__(self, faces, image_size=256, K=None, device="cuda:0"):
        dev = torch.device(device)
        self.faces = torch.as_tensor(np.asarray(faces)).to(torch.int32).contiguous().to(dev)
        self.skip_walks = True            # backward: skip border walks that hold no contributing pixel (prefix counts; same result bit for bit)
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


def make_bbox_square(bbox_xywh: np.ndarray, bbox_expansion: float = 0.0) -> np.ndarray:
    """recon/bbox.py:25-46: square boxes (xywh) around the centres, side = max(w, h) * (1 + expansion)."""
    b = np.array(bbox_xywh, dtype=float).reshape(-1, 4)
    center = np.stack((b[:, 0] + b[:, 2] / 2, b[:, 1] + b[:, 3] / 2), 1)
    side = np.maximum(b[:, 2], b[:, 3])[:, None] * (1 + bbox_expansion)
    return np.hstack((center - side / 2, side, side)).reshape(np.shape(bbox_xywh))


def to_original_bbox(bbox_square: np.ndarray, scale: float, trans: np.ndarray, crop_size: float = 1200) -> np.ndarray:
    """obj_pose_roi.py:110-120: an xywh box in network-input pixels -> original-image pixels (scale = crop_size / net_input_size, trans = the
    crop centre)."""
    out = np.array(bbox_square, dtype=float) * scale
    out[:2] += np.asarray(trans, dtype=float) - crop_size / 2.0
    return out


def cvt_masks(person_mask: torch.Tensor, obj_mask: torch.Tensor) -> torch.Tensor:
    """obj_pose_roi.py:157-170 (PHOSA's occlusion-aware convention): False only where the person covers the pixel and the object does not."""
    fore, ps = obj_mask > 0.5, person_mask > 0.5
    inv = -ps.float()
    inv[fore] = 1.0
    return inv >= 0


def mask_bboxes_xyxy(masks: torch.Tensor) -> np.ndarray:
    """``SilLossROI.masks2bboxes`` (obj_pose_roi.py:173-181 over recon/opt_utils.py:144-155): per mask the bounding box (x1, y1, x2, y2; max
    exclusive) of ``uint8(mask * 255) > 127``; cv2's contour route reduces to the box of the thresholded pixels.  Empty masks keep the
    reference's sentinels."""
    out = []
    for m in masks:
        fg = (m.detach().float() * 255).to(torch.uint8) > 127
        ys, xs = torch.nonzero(fg.any(1)).flatten(), torch.nonzero(fg.any(0)).flatten()
        out.append([50000, 50000, -100, -100] if xs.numel() == 0 else [int(xs[0]), int(ys[0]), int(xs[-1]) + 1, int(ys[-1]) + 1])
    return np.asarray(out)


def roi_setup(person_masks, obj_masks, crop_centers, rend_size=256, bbox_expansion=0.3, camera_params=None, crop_size=1200, net_input_size=512):
    """Everything ``SilLossROI.__init__`` derives from the batch (obj_pose_roi.py:41-70): object-mask boxes -> expanded squares -> both masks
    cropped and resized to ``rend_size`` -> keep mask, reference silhouette, ROI intrinsics.  The crop is detectron2's
    ``BitMasks.crop_and_resize`` = aligned RoIAlign (adaptive sampling) of the boolean mask thresholded at 0.5; detectron2 is not installable
    here, so it is restated on ``torchvision.ops.roi_align`` -- the operator detectron2's ROIAlign wraps (set-up cost once per batch, not on
    the loop).  Returns (keep_mask [B,S,S] float, image_ref [B,S,S] float, K_roi [B,3,3])."""
    from torchvision.ops import roi_align
    pm, om = torch.as_tensor(person_masks), torch.as_tensor(obj_masks)
    B, dev = om.shape[0], om.device
    xyxy = mask_bboxes_xyxy(om).astype(float)
    xywh = np.concatenate([xyxy[:, :2], xyxy[:, 2:] - xyxy[:, :2]], 1)                  # BoxMode XYXY_ABS -> XYWH_ABS
    squares = make_bbox_square(xywh, bbox_expansion)
    sq_xyxy = np.concatenate([squares[:, :2], squares[:, :2] + squares[:, 2:]], 1)      # XYWH_ABS -> XYXY_ABS
    rois = torch.cat([torch.arange(B, dtype=torch.float32)[:, None], torch.as_tensor(sq_xyxy, dtype=torch.float32)], 1).to(dev)
    crop = lambda m: roi_align((m != 0).float()[:, None], rois, (rend_size, rend_size), 1.0, 0, True)[:, 0] >= 0.5
    obj_c, ps_c = crop(om), crop(pm)
    keep = torch.stack([cvt_masks(p, o) for p, o in zip(ps_c, obj_c)]).float()
    ref = (obj_c > 0).float()
    cc = torch.as_tensor(crop_centers).detach().cpu().numpy()
    K = torch.stack([SilLossROI.compute_K_roi(to_original_bbox(sq, crop_size / net_input_size, c, crop_size), **(camera_params or {}))
                     for sq, c in zip(squares, cc)])
    return keep, ref, K


class SilLossROI:
    """Occlusion-aware silhouette loss of recon/obj_pose_roi.py on already-cropped ROI masks.

    keep_mask / image_ref: [B, S, S] as ``cvt_masks`` / ``(obj > 0)`` produce them (obj_pose_roi.py:62-75,157-170);
    K_roi: [B, 3, 3] from ``compute_K_roi`` (:123-155); vertices [V, 3] / faces [F, 3] of the centred template."""

    def __init__(self, keep_mask, image_ref, K_roi, vertices, faces, rend_size=256, device="cuda:0"):
        dev = torch.device(device)
        self.keep_mask, self.image_ref = keep_mask.float().to(dev), image_ref.float().to(dev)
        self.vertices = torch.as_tensor(np.asarray(vertices), dtype=torch.float32).to(dev)
        self.renderer = SilhouetteRenderer(faces, rend_size, K_roi, dev)

    @classmethod
    def from_masks(cls, person_masks, obj_masks, vertices, faces, crop_centers, rend_size=256, bbox_expansion=0.3, device="cuda:0",
                   camera_params=None, crop_size=1200, net_input_size=512):
        """The reference constructor ``SilLossROI(person_masks, obj_masks, temp_mesh, crop_centers, ...)`` (obj_pose_roi.py:21-75): the masks
        are channels 3 / 4 of the network input; see ``roi_setup`` for the steps."""
        keep, ref, K = roi_setup(person_masks, obj_masks, crop_centers, rend_size, bbox_expansion, camera_params, crop_size, net_input_size)
        return cls(keep, ref, K, vertices, faces, rend_size=rend_size, device=device)

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
