"""CPU restatement of the test-time frame preparation -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows the reference's
  * ``BaseDataset.masks2bbox`` / ``center_from_masks`` (data/base_data.py:139-171): uint8 sum of the two masks (wraps like numpy's ``+=``),
    threshold 127, bounding box over all contours == bounding box of the foreground pixels, x + w / y + h exclusive, centre = (min + max) // 2,
  * ``BaseDataset.crop`` (data/base_data.py:204-232): square crop around the centre, zero padding -- including its quirk that a crop running
    past the right / bottom border also drops the image's last column / row (``x2 = min(w - 1, ...)``),
  * ``BehaveDataset.prepare_image_crop`` (data/train_data.py:143-162) and ``BaseDataset.compose_images`` (data/base_data.py:252-265):
    crop -> resize to the network input -> / 255 -> RGB masked by (person > .5) | (object > .5) -> [rgb, person, object] channels first,
  * ``TestDataTriplane.get_item`` (data/testdata_triplane.py:42-74): + the three triplane channels ``png / 255``.
``crop`` / ``compose_images`` / ``prepare_image_crop`` are pinned by tests/golden/frameio_small.npz (the reference's own methods, called
unbound).  PARITY UNPINNED for ``resize_linear_u8``: the reference calls ``cv2.resize(img, size, interpolation=cv2.INTER_LINEAR)`` on uint8
arrays and OpenCV is not installed here (nor vendored by the reference); the function restates OpenCV's portable C path for 8-bit linear
resize as published (modules/imgproc/src/resize.cpp: pixel-centre mapping ``(d + .5) * scale - .5``, 11-bit fixed-point coefficients
rounded with cvRound, horizontal pass in int32, vertical pass ``(((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2``).  Builds of
OpenCV that dispatch to IPP may differ from it by one grey level.
"""
from __future__ import annotations

import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def masks2bbox(masks, thres: int = 127):
    comb = np.zeros_like(masks[0])
    for m in masks:
        comb += m                                                                  # uint8 arithmetic wraps, as in the reference
    comb = np.clip(comb, 0, 255)
    ys, xs = np.nonzero(comb > thres)
    if xs.size == 0:
        return np.array([50000, 50000]), np.array([-100, -100])
    return np.array([xs.min(), ys.min()]), np.array([xs.max() + 1, ys.max() + 1])


def center_from_masks(obj_mask, person_mask):
    bmin, bmax = masks2bbox([person_mask, obj_mask])
    c = (bmin + bmax) // 2
    assert np.sum(c > 0) == 2, "invalid bbox found"
    return c


def crop(img: np.ndarray, center, crop_size) -> np.ndarray:
    h, w = img.shape[:2]
    tl = np.round(np.asarray(center) - np.asarray(crop_size) / 2).astype(int)
    br = np.round(np.asarray(center) + np.asarray(crop_size) / 2).astype(int)
    x1, y1, x2, y2 = max(0, tl[0]), max(0, tl[1]), min(w - 1, br[0]), min(h - 1, br[1])
    out = np.zeros((br[1] - tl[1], br[0] - tl[0]) + img.shape[2:], img.dtype)
    oy, ox = max(0, -tl[1]), max(0, -tl[0])
    out[oy:oy + (y2 - y1), ox:ox + (x2 - x1)] = img[y1:y2, x1:x2]
    return out


def _coeffs(dsize: int, ssize: int):
    """cv::resize INTER_LINEAR tables: source index and the two 11-bit weights per destination index (float32 arithmetic, cvRound)."""
    scale = np.float64(ssize) / dsize
    idx, a0, a1 = np.zeros(dsize, np.int64), np.zeros(dsize, np.int64), np.zeros(dsize, np.int64)
    for d in range(dsize):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - s)
        if s < 0:
            s, f = 0, np.float32(0)
        if s >= ssize - 1:
            s, f = ssize - 1, np.float32(0)
        idx[d] = s
        a0[d] = int(np.rint(np.float32((np.float32(1) - f) * np.float32(COEF_SCALE))))
        a1[d] = int(np.rint(np.float32(f * np.float32(COEF_SCALE))))
    return idx, a0, a1


def resize_linear_u8(img: np.ndarray, dsize) -> np.ndarray:
    """img [H, W] or [H, W, C] uint8 -> [dsize[1], dsize[0](, C)] uint8; dsize = (width, height) as cv2 takes it."""
    assert img.dtype == np.uint8
    dw, dh = int(dsize[0]), int(dsize[1])
    sh, sw = img.shape[:2]
    xi, xa0, xa1 = _coeffs(dw, sw)
    yi, yb0, yb1 = _coeffs(dh, sh)
    src = img.astype(np.int64)
    x1 = np.minimum(xi + 1, sw - 1)
    shp = (1, dw) + (1,) * (img.ndim - 2)
    rows = src[:, xi] * xa0.reshape(shp) + src[:, x1] * xa1.reshape(shp)         # horizontal pass, int32 range
    y1 = np.minimum(yi + 1, sh - 1)
    shp = (dh,) + (1,) * (img.ndim - 1)
    r0, r1 = rows[yi], rows[y1]
    out = (((yb0.reshape(shp) * (r0 >> 4)) >> 16) + ((yb1.reshape(shp) * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def compose_images(obj_mask, person_mask, rgb):
    comb = (person_mask > 0.5) | (obj_mask > 0.5)
    return np.dstack((rgb * np.expand_dims(comb, -1), person_mask, obj_mask))


def prepare_image_crop(rgb, person_mask, obj_mask, crop_size: int = 1200, net_size: int = 512, crop_center=None):
    """-> (images [5, S, S] float32, crop_center [2] int)."""
    c = center_from_masks(obj_mask, person_mask) if crop_center is None else np.asarray(crop_center)
    cs = np.array([crop_size, crop_size])
    r = resize_linear_u8(crop(rgb, c, cs), (net_size, net_size)) / 255.
    p = resize_linear_u8(crop(person_mask, c, cs), (net_size, net_size)) / 255.
    o = resize_linear_u8(crop(obj_mask, c, cs), (net_size, net_size)) / 255.
    return compose_images(o, p, r).transpose((2, 0, 1)).astype(np.float32), c


def test_item(rgb, person_mask, obj_mask, triplane_u8, **kw):
    """``TestDataTriplane.get_item``'s image tensor: [8, S, S] float32 = prepare_image_crop + triplane [S, S, 3] uint8 / 255, and the centre."""
    images, c = prepare_image_crop(rgb, person_mask, obj_mask, **kw)
    tri = (triplane_u8 / 255.).transpose((2, 0, 1))
    return np.concatenate([images, tri], 0).astype(np.float32), c.astype(np.float32)
