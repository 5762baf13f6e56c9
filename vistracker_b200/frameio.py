"""Test-time frame preparation on B200: the numeric part of ``TestDataTriplane.get_item``
(data/testdata_triplane.py:42-74) -- crop centre from the masks, 1200^2 crop, resize to 512^2, /255, background masking, channel stacking --
for a whole batch of decoded uint8 frames in device memory, instead of per frame on DataLoader workers.  Decoding (jpeg / png -> uint8) and
``body_center`` (joint 8 of the SMPL-T fit, ``get_smpl_center``) stay with the caller: the fitting stage already holds the latter on the
device.  OpenCV's resize / contour functions are not available offline; DESIGN.md section 5 says what is pinned and what is not.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib

P, S = _lib.ptr, _lib.stream_ptr
_TABLES = {}


def resize_table(dsize: int, ssize: int) -> np.ndarray:
    """[3, dsize] int32 for cv::resize INTER_LINEAR on 8-bit images: source index, weight of it, weight of the next pixel, the weights in
    11-bit fixed point (``cvRound((1 - f) * 2048)``, ``cvRound(f * 2048)`` with float32 ``f``); pixel centres map as ``(d + .5) * scale - .5``
    and indices past either end are clamped with the fraction dropped."""
    key = (dsize, ssize)
    if key not in _TABLES:
        d = np.arange(dsize, dtype=np.float64)
        f = ((d + 0.5) * (np.float64(ssize) / dsize) - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int64)
        f = (f - s.astype(np.float32)).astype(np.float32)
        lo, hi = s < 0, s >= ssize - 1
        s = np.where(lo, 0, np.where(hi, ssize - 1, s))
        f = np.where(lo | hi, np.float32(0), f).astype(np.float32)
        w1 = np.rint(f * np.float32(2048)).astype(np.int32)
        w0 = np.rint((np.float32(1) - f) * np.float32(2048)).astype(np.int32)
        _TABLES[key] = np.stack([s.astype(np.int32), w0, w1]).astype(np.int32)
    return _TABLES[key]


def crop_center_from_masks(person: torch.Tensor, obj: torch.Tensor, check: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """person / obj [B, H, W] uint8 on the device -> (crop_center [B, 2] float32 (x, y), bbox [B, 4] int32).  ``check`` performs the
    reference's assertions (one device->host copy of B x 2 floats)."""
    B, H, W = person.shape
    person, obj = person.contiguous(), obj.contiguous()
    if person.dtype != torch.uint8 or obj.dtype != torch.uint8 or obj.shape != person.shape:
        raise ValueError("masks must be uint8 tensors of one shape")
    bbox = torch.empty(B, 4, dtype=torch.int32, device=person.device)
    center = torch.empty(B, 2, dtype=torch.float32, device=person.device)
    with torch.cuda.device(person.device):
        _lib.call("vt_mask_bbox", P(person), P(obj), B, H, W, 127, P(bbox), P(center), S())
    if check:
        c = center.cpu()
        if not bool(((c > 0).sum(1) == 2).all()) or not bool(((c[:, 0] < W) & (c[:, 1] < W)).all()):                # base_data.py:163-168 (sic: iw twice)
            raise AssertionError(f"invalid bbox / crop center found: {c.tolist()}")
    return center, bbox


def prepare_images(rgb: torch.Tensor, person: torch.Tensor, obj: torch.Tensor, triplane: Optional[torch.Tensor] = None,
                   crop_center: Optional[torch.Tensor] = None, crop_size: int = 1200, net_size: int = 512, check: bool = True):
    """rgb [B, H, W, 3], person / obj [B, H, W], triplane [B, S, S, 3] (all uint8, device) -> (images [B, 8 | 5, S, S] float32, crop_center
    [B, 2] float32): the 'images' and 'crop_center' entries of the reference's batch dict."""
    B, H, W = person.shape
    dev = rgb.device
    if rgb.dtype != torch.uint8 or tuple(rgb.shape) != (B, H, W, 3):
        raise ValueError(f"rgb must be uint8 [B, H, W, 3], got {rgb.dtype} {tuple(rgb.shape)}")
    if triplane is not None and (triplane.dtype != torch.uint8 or tuple(triplane.shape) != (B, net_size, net_size, 3)):
        raise ValueError(f"triplane must be uint8 [B, {net_size}, {net_size}, 3]")
    if crop_center is None:
        crop_center, _ = crop_center_from_masks(person, obj, check)
    crop_center = crop_center.to(dev, torch.float32).contiguous()
    C = 8 if triplane is not None else 5
    images = torch.empty(B, C, net_size, net_size, device=dev)
    tab = torch.from_numpy(resize_table(net_size, crop_size)).to(dev)
    with torch.cuda.device(dev):
        _lib.call("vt_prepare_image_crop", P(rgb.contiguous()), P(person.contiguous()), P(obj.contiguous()),
                  P(triplane.contiguous()) if triplane is not None else None, B, H, W, P(crop_center), crop_size, net_size, P(tab), P(images), C, S())
    return images, crop_center
