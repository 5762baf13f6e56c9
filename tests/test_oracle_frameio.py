"""oracle/frameio_ref.py against the reference's own crop / compose / resize / masks2bbox / prepare_image_crop run with the real OpenCV 4.13
(tests/golden/frameio_small.npz, written by tests/golden/make_golden.py --only frameio) -- CPU only.  The restatements of cv2.resize
(INTER_LINEAR, uint8) and of threshold + findContours + boundingRect are pinned bit for bit to cv2's own outputs."""
import hashlib
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


def test_resize_restatement_equals_cv2_bit_for_bit():
    """cv2.resize(img, dsize, interpolation=cv2.INTER_LINEAR) on uint8 (data/base_data.py:249): production ratio 1200 -> 512 on 3 and 1
    channels, non-square, identity and an up-scale; the full-size 1200 x 1200 -> 512 x 512 results through their SHA-256."""
    g = np.load(GOLD)
    for tag, dsize in (("r300_128_rgb", (128, 128)), ("r300_128_mask", (128, 128)), ("r300x200_128x96", (128, 96)), ("r75_32", (32, 32)),
                       ("r64_64", (64, 64))):
        out = FR.resize_linear_u8(g[f"{tag}_src"], dsize)
        assert out.dtype == np.uint8 and out.shape == g[f"{tag}_dst"].shape, tag
        assert np.array_equal(out, g[f"{tag}_dst"]), (tag, int((out != g[f"{tag}_dst"]).sum()))
    # enlarging (37 -> 100) is NOT on the reference's path (it only shrinks the 1200-pixel crop to 512): this cv2 build takes another code
    # path there and differs from the portable fixed-point one by one grey level on ~0.25 % of the pixels
    up = FR.resize_linear_u8(g["r37_100_src"], (100, 100)).astype(np.int32) - g["r37_100_dst"].astype(np.int32)
    assert int(np.abs(up).max()) <= 1 and float((up != 0).mean()) < 0.01
    rgb, person, _ = synthetic_camera_frame(1200, 1200, seed=41)
    assert hashlib.sha256(FR.resize_linear_u8(rgb, (512, 512)).tobytes()).hexdigest() == str(g["r1200_512_rgb_sha256"])
    assert hashlib.sha256(FR.resize_linear_u8(person, (512, 512)).tobytes()).hexdigest() == str(g["r1200_512_mask_sha256"])


def test_masks2bbox_restatement_equals_cv2_contours():
    """cv2.threshold + findContours + boundingRect (data/base_data.py:139-157) on masks with speckle at / below / above the threshold, a
    uint8 wrap-around of the mask sum and the three synthetic camera frames."""
    g = np.load(GOLD)
    for k in range(4):
        lo, hi = FR.masks2bbox([g[f"bbox{k}_a"], g[f"bbox{k}_b"]])
        assert np.concatenate([lo, hi]).tolist() == g["bbox_ref"][k].tolist(), k
    H, W = int(g["H"]), int(g["W"])
    for i, (tag, ctr) in enumerate(CENTERS.items()):
        _, person, obj = synthetic_camera_frame(H, W, seed=30 + i, center=ctr)
        lo, hi = FR.masks2bbox([person, obj])
        assert np.concatenate([lo, hi]).tolist() == g[f"{tag}_bbox"].tolist(), tag
