"""Experiment options for the SIF-Net hot path.

The reference reads a JSON-with-``//``-comments file into an ``argparse.Namespace``
(``config/config_loader.py:24-45``) and the model classes pick fields out of it with
``'key' in opt`` tests (``model/chore.py:35-96``, ``model/chore_triplane.py:23-45``,
``model/HGFilters.py:59-160``).  This module accepts the same Namespace / JSON and resolves it into
the fixed set of numbers the B200 kernels need (:class:`SIFNetDims`).  Only the fields that change
the arithmetic of the hot path are kept; training-only fields are carried through untouched.
"""
from __future__ import annotations

import json
from argparse import Namespace
from collections import OrderedDict
from dataclasses import dataclass
from os.path import join

# The arithmetic-relevant part of config/tri-vis-l2.json (reference file, lines 45-80).
TRI_VIS_L2 = OrderedDict(
    exp_name="tri-vis-l2",
    model_name="chore-triplane-vis",
    net_img_size=[512, 512],
    loadSize=1200,
    gpu_id=0,
    z_0=2.2,
    input_type="RGBM3",
    norm="group",
    num_stack=3,
    num_hourglass=2,
    skip_hourglass=True,
    hg_down="ave_pool",
    hourglass_dim=256,
    z_feat="smpl-triplane",
    vis_activation="sigmoid",
    vis_loss="l2",
    loss_weights=[1.0, 1.0, 0.006, 500, 1000, 1000],
    triplane_encoder_stack=3,
    triplane_shared_encoder=True,
    triplane_hg_dim=64,
    triplane_tmpx_dim=32,
    projection_mode="perspective",
    filter_val=0.004,
    sparse_thres=0.03,
)


def default_options() -> Namespace:
    """Namespace equivalent to ``load_configs('tri-vis-l2')`` for the fields this path reads."""
    ns = Namespace()
    ns.__dict__ = OrderedDict(TRI_VIS_L2)
    return ns


def load_configs(exp_name: str, configs_dir: str = "config") -> Namespace:
    """Same contract as the reference loader (``config/config_loader.py:24-45``): read
    ``<configs_dir>/<exp_name>.json``, strip ``//`` comments, keep key order, and check that
    ``camera_params['crop_size']`` agrees with ``loadSize`` when both are present."""
    text = []
    with open(join(configs_dir, exp_name + ".json"), "r") as f:
        for line in f:
            text.append(line.split("//")[0])
    ns = Namespace()
    ns.__dict__ = json.loads("\n".join(text), object_pairs_hook=OrderedDict)
    if "camera_params" in ns and "loadSize" in ns:
        if ns.camera_params["crop_size"] != ns.loadSize:
            raise AssertionError("please check camera params and crop size!")
    return ns


# Input channel count per ``input_type`` (model/HGFilters.py:63-113); the stacked variants need
# ``frame_deltas`` and are outside this path.
_INPUT_CHANNELS = {
    "RGB": 3, "RGBD": 4, "RGBN": 5, "RGBM2": 5, "RGBM3": 5, "RGBM4": 5, "RGBMD": 6, "RGBMD2": 6,
    "RGBM": 4, "RGBMN": 8, "mask": 1, "mask2": 2,
}

# BEHAVE Kinect colour camera, model/camera.py:26-40.  NB: fy_px = fy * image_WIDTH there.
KINECT_FX_PX = 979.7844
KINECT_FY_PX = 979.840
KINECT_CX_PX = 1018.952
KINECT_CY_PX = 779.486


@dataclass(frozen=True)
class EncoderDims:
    """One stacked-hourglass encoder (model/HGFilters.py:56-160)."""
    in_ch: int
    stem_ch: int      # ``tmpx_dim``: channels of the 7x7/s2 stem and of the returned ``tmpx``
    out_ch: int       # ``hourglass_dim``: channels of the ``l{i}`` heads
    num_stack: int
    depth: int        # ``num_hourglass``
    feat_ch: int = 256


@dataclass(frozen=True)
class SIFNetDims:
    rgb: EncoderDims
    tri: EncoderDims
    hidden: int
    num_parts: int
    feature_size: int
    crop_size: float
    fx_px: float
    fy_px: float
    cx_px: float
    cy_px: float
    z0: float = 2.2          # hard-coded in model/chore_triplane.py:216
    out_dist: float = 5.0    # model/chore.py:93


def _get(opt, key, default=None):
    return getattr(opt, key) if key in opt else default


def resolve_dims(opt: Namespace, num_parts: int = 14) -> SIFNetDims:
    """Turn the reference's ``opt`` Namespace into kernel dimensions, rejecting configurations the
    B200 path does not implement (it is a drop-in for ``tri-vis-l2``-shaped models only)."""
    if _get(opt, "z_feat") != "smpl-triplane":
        raise NotImplementedError(f"z_feat={_get(opt, 'z_feat')!r}: only 'smpl-triplane' is built")
    if _get(opt, "norm") != "group":
        raise NotImplementedError("only GroupNorm encoders are built (norm='group')")
    if _get(opt, "hg_down") != "ave_pool":
        raise NotImplementedError("only hg_down='ave_pool' is built")
    if not _get(opt, "skip_hourglass", False):
        raise AssertionError("skip_hourglass must be true (model/chore_triplane.py:126)")
    if not _get(opt, "triplane_shared_encoder", False):
        raise NotImplementedError("only the shared triplane encoder is built")
    if _get(opt, "vis_activation") != "sigmoid":
        raise NotImplementedError("vis_activation must be 'sigmoid' (model/chore_tri_vis.py:23-26)")
    if _get(opt, "vis_loss") not in ("l1", "l2"):
        raise AssertionError("vis_loss must be l1 or l2 (model/chore_tri_vis.py:29)")
    in_type = _get(opt, "input_type")
    if in_type not in _INPUT_CHANNELS:
        raise ValueError(f"invalid input specification: {in_type}")
    if _INPUT_CHANNELS[in_type] != 5:
        raise NotImplementedError("CHORETriplane.filter feeds images[:, :5] to the RGB encoder")
    rgb = EncoderDims(in_ch=5, stem_ch=int(_get(opt, "tmpx_dim", 64)),
                      out_ch=int(opt.hourglass_dim), num_stack=int(opt.num_stack),
                      depth=int(opt.num_hourglass))
    tri = EncoderDims(in_ch=1, stem_ch=int(opt.triplane_tmpx_dim), out_ch=int(opt.triplane_hg_dim),
                      num_stack=int(opt.triplane_encoder_stack), depth=int(opt.num_hourglass))
    hidden = int(_get(opt, "hidden_dim", 128))
    # model/chore.py:57-62 + chore_triplane.py:53-58.  NB the reference sizes the skip feature as
    # hourglass_dim // 4 (= 64 = the default stem width); a non-default tmpx_dim would break it there too.
    feature_size = rgb.out_ch + 3 + (tri.out_ch + tri.stem_ch) * 3 + rgb.out_ch // 4
    if rgb.stem_ch != rgb.out_ch // 4:
        raise NotImplementedError("tmpx_dim must equal hourglass_dim // 4 (model/chore.py:61)")
    cam = _get(opt, "camera_params")
    if cam is None:
        crop, fx, fy, cx, cy = float(opt.loadSize), KINECT_FX_PX, KINECT_FY_PX, KINECT_CX_PX, KINECT_CY_PX
    else:
        width = float(cam.get("image_width", 2048))
        crop = float(cam.get("crop_size", 1200))
        fx = float(cam.get("fx", KINECT_FX_PX / 2048.)) * width
        fy = float(cam.get("fy", KINECT_FY_PX / 2048.)) * width
        cx = float(cam.get("cx", KINECT_CX_PX / 2048.)) * width
        cy = float(cam.get("cy", KINECT_CY_PX / 2048.)) * width
    return SIFNetDims(rgb=rgb, tri=tri, hidden=hidden, num_parts=num_parts, feature_size=feature_size,
                      crop_size=crop, fx_px=fx, fy_px=fy, cx_px=cx, cy_px=cy)
