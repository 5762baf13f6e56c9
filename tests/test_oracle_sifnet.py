"""The CPU oracle against golden vectors produced by the unmodified reference (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import torch

from conftest import GOLDEN, rel_err
from oracle import sifnet_ref as R
from vistracker_b200.config import default_options, resolve_dims
from vistracker_b200.synth import sifnet_spec, synthetic_frames, synthetic_state_dict

DIMS = resolve_dims(default_options())
CAM = (DIMS.fx_px, DIMS.fy_px, DIMS.cx_px, DIMS.cy_px, DIMS.crop_size)


def test_state_dict_spec_matches_reference_keys():
    with open(os.path.join(GOLDEN, "sifnet_keys.json")) as f:
        ref = [(k, tuple(s)) for k, s in json.load(f)]
    ours = [(k, tuple(s)) for k, s, _ in sifnet_spec(DIMS)]
    assert len(ref) == 706
    assert ours == ref
    sd = synthetic_state_dict(DIMS, seed=0)
    assert sum(v.numel() for k, v in sd.items() if ".downsample.0." not in k) == 22093661   # generator.py:7-10


def _check(maps, outs, g, tol):
    for k in ("im_feat", "tmpx"):
        assert rel_err(maps[k], g[k]) < tol, k
    for v in range(3):
        assert rel_err(maps["tri_feat"][v], g[f"tri_feat{v}"]) < tol
        assert rel_err(maps["tri_tmpx"][v], g[f"tri_tmpx{v}"]) < tol
    for name, o in zip(("df", "pca", "parts", "centers", "vis"), outs):
        assert o.shape == g[name].shape
        assert rel_err(o.detach(), g[name]) < tol, name


def test_oracle_small_matches_reference(golden):
    g = golden("sifnet_small.npz")
    sd = synthetic_state_dict(DIMS, seed=0)
    images, points, crop, body = synthetic_frames(2, size=64, seed=11, n_points=301, jitter=True)
    with torch.no_grad():
        maps = R.sif_filter(sd, images)
        feat, xy = R.point_features(maps, points, crop, body, CAM)
    assert rel_err(xy, g["xy"]) < 1e-6
    assert rel_err(feat, g["features"]) < 1e-5
    assert not ((np.abs(g["xy"]) <= 1).all()), "fixture must contain out-of-image points"
    for name, idx in (("grad_h", 0), ("grad_o", 1)):
        pts = points.clone().requires_grad_(True)
        outs = R.sif_query(sd, maps, pts, crop, body, CAM)
        torch.clamp(outs[0][:, idx], max=2.0).sum().backward()
        assert rel_err(pts.grad, g[name]) < 1e-4, name
    _check(maps, outs, g, 1e-5)
    out_of_img = (np.abs(g["xy"]) > 1).any(1)
    assert (outs[0].detach().numpy().transpose(0, 2, 1)[out_of_img] == 5.0).all()


def test_oracle_config1_matches_reference(golden):
    g = golden("sifnet_c1.npz")
    sd = synthetic_state_dict(DIMS, seed=0)
    images, points, crop, body = synthetic_frames(1, size=512, seed=0, n_points=2000)
    with torch.no_grad():
        maps = R.sif_filter(sd, images)
        outs = R.sif_query(sd, maps, points, crop, body, CAM)
    sub = {k: (v[:, :, ::8, ::8] if torch.is_tensor(v) else [t[:, :, ::8, ::8] for t in v]) for k, v in maps.items()}
    _check(sub, outs, g, 1e-5)
