"""oracle/frameio_ref.py against the reference's own crop / compose / prepare_image_crop (tests/golden/frameio_small.npz) -- CPU only.
cv2.resize and cv2.findContours are unpinned (OpenCV is not installed; see the oracle's header): their restatements are checked against
independent formulations here."""
import os

import numpy as np

from oracle import frameio_ref as FR
from vistracker_b200.synth import synthetic_camera_frame

GOLD = os.path.join(os.path.dirname(__file__), "golden", "frameio_small.npz")
CENTERS = {"mid": None, "right_bottom": (205.0, 150.0), "top_left": (30.0, 25.0)}


def test_crop_and_pipeline_match_reference_methods():
    g = np.load(GOLD)
    H, W, CROP, NET = int(g["H"]), int(g["W"]), int(g["crop"]), int(g["net"])
    for i, (tag, ctr) in enumerate(CENTERS.items()):
        rgb, person, obj = synthetic_camera_frame(H, W, seed=30 + i, center=ctr)
        c = FR.center_from_masks(obj, person)
        assert np.array_equal(c, g[f"{tag}_center"])
        assert np.array_equal(FR.crop(rgb, c, np.array([CROP, CROP])), g[f"{tag}_crop_rgb"])
        images, c2 = FR.prepare_image_crop(rgb, person, obj, CROP, NET)
        assert images.dtype == np.float32 and np.array_equal(images, g[f"{tag}_images"])
    # the crops past the borders really are padded (and the reference's dropped last column is reproduced)
    assert (g["right_bottom_crop_rgb"][:, -1] == 0).all() and (g["top_left_crop_rgb"][0] == 0).all()


def test_resize_restatement_properties():
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (150, 150, 3)).astype(np.uint8)
    out = FR.resize_linear_u8(img, (64, 64))
    # float bilinear with pixel-centre mapping: the fixed-point path stays within one grey level of it
    s = 150 / 64
    fx = (np.arange(64) + 0.5) * s - 0.5
    x0 = np.floor(fx).astype(int); wx = fx - x0
    x0c, x1c = np.clip(x0, 0, 149), np.clip(x0 + 1, 0, 149)
    f = img.astype(np.float64)
    rows = f[:, x0c] * (1 - wx)[None, :, None] + f[:, x1c] * wx[None, :, None]
    ref = rows[x0c] * (1 - wx)[:, None, None] + rows[x1c] * wx[:, None, None]
    assert np.abs(out.astype(np.float64) - ref).max() <= 1.0
    assert np.array_equal(FR.resize_linear_u8(np.full((150, 150), 200, np.uint8), (64, 64)), np.full((64, 64), 200, np.uint8))
    assert np.array_equal(FR.resize_linear_u8(img, (150, 150)), img)                # identity scale
    # bbox: wrap-around of the uint8 sum (128 + 128 = 0) is reproduced, x + w is exclusive
    a, b = np.zeros((20, 30), np.uint8), np.zeros((20, 30), np.uint8)
    a[5:9, 10:14] = 255; b[7:12, 12:20] = 255; a[15, 25] = 128; b[15, 25] = 128
    lo, hi = FR.masks2bbox([a, b])
    assert lo.tolist() == [10, 5] and hi.tolist() == [20, 12]
