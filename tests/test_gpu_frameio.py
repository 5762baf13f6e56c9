"""vt_mask_bbox / vt_prepare_image_crop through vistracker_b200.frameio against the reference's prepare_image_crop outputs
(tests/golden/frameio_small.npz) and the numpy restatement at the real frame size (2048 x 1536 -> 1200^2 -> 512^2): bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import frameio_ref as FR
from vistracker_b200.synth import synthetic_camera_frame

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "frameio_small.npz")
CENTERS = {"mid": None, "right_bottom": (205.0, 150.0), "top_left": (30.0, 25.0)}


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def test_small_frames_match_reference_golden():
    _need_gpu()
    from vistracker_b200.frameio import crop_center_from_masks, prepare_images
    g = np.load(GOLD)
    H, W, CROP, NET = int(g["H"]), int(g["W"]), int(g["crop"]), int(g["net"])
    frames = [synthetic_camera_frame(H, W, seed=30 + i, center=c) for i, c in enumerate(CENTERS.values())]
    rgb, person, obj = (torch.from_numpy(np.stack([f[k] for f in frames])).cuda() for k in range(3))
    center, bbox = crop_center_from_masks(person, obj)
    images, cc = prepare_images(rgb, person, obj, crop_size=CROP, net_size=NET)
    assert images.shape == (3, 5, NET, NET) and torch.equal(cc, center)
    for i, tag in enumerate(CENTERS):
        assert np.array_equal(center[i].cpu().numpy(), g[f"{tag}_center"].astype(np.float32)), tag
        assert np.array_equal(images[i].cpu().numpy(), g[f"{tag}_images"]), tag
        lo, hi = FR.masks2bbox([frames[i][1], frames[i][2]])
        assert bbox[i].tolist() == [lo[0], lo[1], hi[0], hi[1]]


def test_full_size_frames_with_triplane_channels():
    _need_gpu()
    from vistracker_b200.frameio import prepare_images
    frames = [synthetic_camera_frame(1536, 2048, seed=s, center=c) for s, c in ((1, None), (2, (1750.0, 1100.0)))]
    rgb, person, obj = (torch.from_numpy(np.stack([f[k] for f in frames])).cuda() for k in range(3))
    tri = torch.from_numpy((np.random.default_rng(0).random((2, 512, 512, 3)) > 0.7).astype(np.uint8) * 255).cuda()
    images, center = prepare_images(rgb, person, obj, tri)
    assert images.shape == (2, 8, 512, 512) and images.dtype == torch.float32
    for i, f in enumerate(frames):
        ref, c = FR.test_item(f[0], f[1], f[2], tri[i].cpu().numpy())
        assert np.array_equal(center[i].cpu().numpy(), c)
        assert np.array_equal(images[i].cpu().numpy(), ref)
    # a given centre (the reference reuses 'old_crop_center' when it re-crops) and no triplane
    images5, _ = prepare_images(rgb, person, obj, crop_center=torch.tensor([[900.0, 700.0], [1000.0, 800.0]]))
    ref5, _ = FR.prepare_image_crop(*frames[0], crop_center=np.array([900, 700]))
    assert images5.shape == (2, 5, 512, 512) and np.array_equal(images5[0].cpu().numpy(), ref5)


def test_empty_masks_raise_like_the_reference():
    _need_gpu()
    from vistracker_b200.frameio import crop_center_from_masks
    z = torch.zeros(1, 64, 64, dtype=torch.uint8, device="cuda")
    center, bbox = crop_center_from_masks(z, z, check=False)
    assert bbox[0].tolist() == [50000, 50000, -100, -100]
    with pytest.raises(AssertionError, match="invalid"):
        crop_center_from_masks(z, z)
    with pytest.raises(ValueError, match="uint8"):
        crop_center_from_masks(z.float(), z)
