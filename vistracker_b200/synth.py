"""Seeded synthetic checkpoints and inputs.

The reference ships neither its trained checkpoints nor BEHAVE data (README.md:37,42,62), so parity
and throughput are established on a random-init network with a fixed seed and on synthetic frames
(SURVEY.md section 8(d)).  Everything here is generated with ``numpy.random.Generator(PCG64(seed))`` in a
fixed key order so that the build container (where the golden vectors are produced from the real
reference code) and the GPU box (where only this repo exists) see bit-identical tensors.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import List, Tuple

import numpy as np
import torch

from .config import EncoderDims, SIFNetDims

Spec = List[Tuple[str, Tuple[int, ...], str]]   # (key, shape, kind) kind in {conv_w, bias, gn_w, gn_b}


def _convblock_spec(prefix: str, cin: int, cout: int) -> Spec:
    """Keys of one ``ConvBlock`` in ``state_dict()`` order (model/net_util.py:346-372): three bias-free
    3x3 convs, four GroupNorms (``bn4`` exists even when unused) and, when cin != cout, a ``downsample``
    Sequential that re-lists ``bn4`` as ``downsample.0`` followed by a bias-free 1x1 conv."""
    half, quarter = cout // 2, cout // 4
    s: Spec = [
        (f"{prefix}.conv1.weight", (half, cin, 3, 3), "conv_w"),
        (f"{prefix}.conv2.weight", (quarter, half, 3, 3), "conv_w"),
        (f"{prefix}.conv3.weight", (quarter, quarter, 3, 3), "conv_w"),
    ]
    for name, c in (("bn1", cin), ("bn2", half), ("bn3", quarter), ("bn4", cin)):
        s += [(f"{prefix}.{name}.weight", (c,), "gn_w"), (f"{prefix}.{name}.bias", (c,), "gn_b")]
    if cin != cout:
        s += [(f"{prefix}.downsample.0.weight", (cin,), "alias:" + f"{prefix}.bn4.weight"),
              (f"{prefix}.downsample.0.bias", (cin,), "alias:" + f"{prefix}.bn4.bias"),
              (f"{prefix}.downsample.2.weight", (cout, cin, 1, 1), "conv_w")]
    return s


def _hourglass_spec(prefix: str, level: int, ch: int) -> Spec:
    """``HourGlass._generate_network`` registration order (model/HGFilters.py:14-24)."""
    s = _convblock_spec(f"{prefix}.b1_{level}", ch, ch) + _convblock_spec(f"{prefix}.b2_{level}", ch, ch)
    if level > 1:
        s += _hourglass_spec(prefix, level - 1, ch)
    else:
        s += _convblock_spec(f"{prefix}.b2_plus_{level}", ch, ch)
    s += _convblock_spec(f"{prefix}.b3_{level}", ch, ch)
    return s


def encoder_spec(prefix: str, d: EncoderDims) -> Spec:
    """``HGFilter.__init__`` registration order (model/HGFilters.py:118-160)."""
    f = d.feat_ch
    s: Spec = [(f"{prefix}.conv1.weight", (d.stem_ch, d.in_ch, 7, 7), "conv_w"),
               (f"{prefix}.conv1.bias", (d.stem_ch,), "bias"),
               (f"{prefix}.bn1.weight", (d.stem_ch,), "gn_w"), (f"{prefix}.bn1.bias", (d.stem_ch,), "gn_b")]
    s += _convblock_spec(f"{prefix}.conv2", d.stem_ch, 128)
    s += _convblock_spec(f"{prefix}.conv3", 128, 128)
    s += _convblock_spec(f"{prefix}.conv4", 128, f)
    for i in range(d.num_stack):
        s += _hourglass_spec(f"{prefix}.m{i}", d.depth, f)
        s += _convblock_spec(f"{prefix}.top_m_{i}", f, f)
        s += [(f"{prefix}.conv_last{i}.weight", (f, f, 1, 1), "conv_w"), (f"{prefix}.conv_last{i}.bias", (f,), "bias"),
              (f"{prefix}.bn_end{i}.weight", (f,), "gn_w"), (f"{prefix}.bn_end{i}.bias", (f,), "gn_b"),
              (f"{prefix}.l{i}.weight", (d.out_ch, f, 1, 1), "conv_w"), (f"{prefix}.l{i}.bias", (d.out_ch,), "bias")]
        if i < d.num_stack - 1:
            s += [(f"{prefix}.bl{i}.weight", (f, f, 1, 1), "conv_w"), (f"{prefix}.bl{i}.bias", (f,), "bias"),
                  (f"{prefix}.al{i}.weight", (f, d.out_ch, 1, 1), "conv_w"), (f"{prefix}.al{i}.bias", (f,), "bias")]
    return s


# decoder heads in module-registration order: CHORE.__init__ creates df, part, pca, center
# (model/chore.py:72-79); init_others re-assigns center_predictor in place and appends visib_predictor
# (model/chore_tri_vis.py:17-28).  Output widths: df 2, parts 14, pca 9, centers 3, visibility 1.
DECODER_HEADS = (("df", 2), ("part_predictor", None), ("pca_predictor", 9), ("center_predictor", 3),
                 ("visib_predictor", 1))


def decoder_spec(dims: SIFNetDims) -> Spec:
    s: Spec = []
    h = dims.hidden
    for name, out in DECODER_HEADS:
        out = dims.num_parts if out is None else out
        for idx, (co, ci) in zip((0, 2, 4, 6), ((h, dims.feature_size), (h, h), (h, h), (out, h))):
            s += [(f"{name}.{idx}.weight", (co, ci, 1), "conv_w"), (f"{name}.{idx}.bias", (co,), "bias")]
    return s


def sifnet_spec(dims: SIFNetDims) -> Spec:
    """All 706 ``CHORETriplaneVisibility.state_dict()`` entries for tri-vis-l2, in order."""
    return encoder_spec("image_filter", dims.rgb) + decoder_spec(dims) + encoder_spec("triplane_encoder", dims.tri)


def synthetic_state_dict(dims: SIFNetDims, seed: int = 0, plain_init: bool = False) -> "OrderedDict[str, torch.Tensor]":
    """Random checkpoint with the reference's key set.

    ``plain_init=True`` reproduces the *distribution* of ``init_weights`` (N(0, 0.02) weights, zero biases,
    identity GroupNorm affine; model/net_util.py:230-244).  The default additionally randomises biases and
    GroupNorm affines so that a kernel that drops one of them cannot pass parity.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, shape, kind in sifnet_spec(dims):
        if kind.startswith("alias:"):
            sd[key] = sd[kind[6:]]
            continue
        if kind == "conv_w":
            a = rng.standard_normal(shape, dtype=np.float32) * np.float32(0.02)
        elif kind == "bias":
            a = np.zeros(shape, np.float32) if plain_init else rng.standard_normal(shape, dtype=np.float32) * np.float32(0.05)
        elif kind == "gn_w":
            a = np.ones(shape, np.float32) if plain_init else (1 + rng.standard_normal(shape, dtype=np.float32) * np.float32(0.1))
        elif kind == "gn_b":
            a = np.zeros(shape, np.float32) if plain_init else rng.standard_normal(shape, dtype=np.float32) * np.float32(0.1)
        else:
            raise AssertionError(kind)
        sd[key] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return sd


def synthetic_frames(batch: int, size: int = 512, seed: int = 0, n_points: int = 2000, jitter: bool = False):
    """Seeded inputs of SURVEY.md section 8(d): ``images [B,8,S,S]`` with channels 3-7 binarised and RGB masked
    by (person OR object); query points uniform in the generator's 2 x 3 x 1.2 m box around the body centre
    (recon/gen/generator_triplane.py:46-53); ``crop_center`` in 2048x1536 pixel space; ``body_center``.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    img = rng.random((batch, 8, size, size), dtype=np.float32)
    img[:, 3:] = (img[:, 3:] > 0.5).astype(np.float32)
    img[:, :3] *= np.maximum(img[:, 3:4], img[:, 4:5])
    crop = np.tile(np.array([[1024.0, 768.0]], np.float32), (batch, 1))
    body = np.tile(np.array([[0.0, 0.0, 2.2]], np.float32), (batch, 1))
    if jitter:
        crop += rng.uniform(-100, 100, (batch, 2)).astype(np.float32)
        body += (rng.uniform(-1, 1, (batch, 3)) * np.array([0.3, 0.2, 0.2])).astype(np.float32)
    pts = rng.random((batch, n_points, 3), dtype=np.float32)
    pts = pts * np.array([2.0, 3.0, 1.2], np.float32) - np.array([1.0, 1.5, 0.6], np.float32) + body[:, None, :]
    return (torch.from_numpy(img), torch.from_numpy(pts.astype(np.float32)), torch.from_numpy(crop),
            torch.from_numpy(body))


def infill_spec(opt) -> list:
    """(key, shape, kind) of a ``ConditionalMInfiller`` checkpoint (model/infill/mfiller_cond.py:20-55; layers of
    model/transformers/former_deci.py:31-51 around nn.MultiheadAttention), in state_dict order."""
    g = (lambda k: opt[k]) if isinstance(opt, dict) else (lambda k: getattr(opt, k))
    spec = [("feat_proj_smpl.weight", (g("d_model_smpl"), g("dim_smpl")), "w"), ("feat_proj_smpl.bias", (g("d_model_smpl"),), "b"),
            ("feat_proj_obj.weight", (g("d_model_obj"), g("dim_obj")), "w"), ("feat_proj_obj.bias", (g("d_model_obj"),), "b")]
    d_joint = g("d_model_smpl") + g("d_model_obj")
    for name, D in (("smpl", g("d_model_smpl")), ("obj", g("d_model_obj")), ("joint", d_joint)):
        F = g("dim_forward_" + name)
        for i in range(g("num_layers_" + name)):
            p = f"encoder_{name}.encoder.layers.{i}."
            spec += [(p + "self_attn.in_proj_weight", (3 * D, D), "w"), (p + "self_attn.in_proj_bias", (3 * D,), "b"),
                     (p + "self_attn.out_proj.weight", (D, D), "w"), (p + "self_attn.out_proj.bias", (D,), "b"),
                     (p + "linear1.weight", (F, D), "w"), (p + "linear1.bias", (F,), "b"),
                     (p + "linear2.weight", (D, F), "w"), (p + "linear2.bias", (D,), "b"),
                     (p + "norm1.weight", (D,), "ln_w"), (p + "norm1.bias", (D,), "ln_b"),
                     (p + "norm2.weight", (D,), "ln_w"), (p + "norm2.bias", (D,), "ln_b")]
        if g("pre_norm_" + name):
            spec += [(f"encoder_{name}.encoder.norm.weight", (D,), "ln_w"), (f"encoder_{name}.encoder.norm.bias", (D,), "ln_b")]
    dims = [d_joint] + list(g("hidden_dims")) + [g("out_dim")]
    for i in range(len(dims) - 1):
        spec += [(f"predictor.{2 * i}.weight", (dims[i + 1], dims[i]), "w"), (f"predictor.{2 * i}.bias", (dims[i + 1],), "b")]
    return spec


def synthetic_infill_state_dict(opt, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Seeded random HVOP-Net checkpoint with the reference's key set: weights N(0, 1.5 / fan_in) (sharp enough that the attention is far
    from uniform), biases N(0, 0.1), LayerNorm affines around (1, 0)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, shape, kind in infill_spec(opt):
        if kind == "w":
            a = rng.standard_normal(shape, dtype=np.float32) * np.float32(np.sqrt(1.5 / shape[1]))
        elif kind == "ln_w":
            a = 1 + rng.standard_normal(shape, dtype=np.float32) * np.float32(0.1)
        else:
            a = rng.standard_normal(shape, dtype=np.float32) * np.float32(0.1)
        sd[key] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return sd


def synthetic_infill_sequence(L: int, seed: int = 0, occluded=((70, 130), (215, 290))):
    """A smooth synthetic trajectory in the in-filler's input format: rot6d_smpl [L,144], trans_smpl [L,3], rot6d_obj [L,6] (noisy inside the
    occluded spans), trans_obj [L,3], occ_ratios [L] (visible fraction, low inside the spans).  float32 numpy."""
    rng = np.random.Generator(np.random.PCG64(seed))
    t = np.arange(L, dtype=np.float64)[:, None] / 30.0
    rot6d_smpl = np.sin(t * rng.uniform(0.3, 2.0, (1, 144)) + rng.uniform(0, 6.28, (1, 144))) * 0.8
    trans_smpl = np.array([[0.1, -0.2, 2.4]]) + 0.3 * np.sin(t * np.array([[0.7, 1.1, 0.4]]))
    rot6d_obj = np.sin(t * rng.uniform(0.2, 1.0, (1, 6)) + rng.uniform(0, 6.28, (1, 6)))
    trans_obj = np.array([[0.3, 0.1, 2.2]]) + 0.2 * np.sin(t * np.array([[0.5, 0.9, 0.3]]))
    occ = 0.75 + 0.2 * rng.random(L)
    for a, b in occluded:
        a, b = min(a, L), min(b, L)
        occ[a:b] = 0.05 + 0.3 * rng.random(b - a)
        rot6d_obj[a:b] += 0.5 * rng.standard_normal((b - a, 6))
    f = lambda x: np.ascontiguousarray(x, dtype=np.float32)
    return f(rot6d_smpl), f(trans_smpl), f(rot6d_obj), f(trans_obj), f(occ)


def synthetic_camera_frame(H: int = 1536, W: int = 2048, seed: int = 0, center=None, extent=(0.22, 0.42)):
    """A seeded stand-in for one Kinect frame: rgb [H,W,3] uint8 noise + gradients, a person mask (ellipse) and an object mask (rotated box)
    as uint8 0/255 with soft (anti-aliased, jpeg-like) borders, placed around ``center`` (x, y) -- default image centre."""
    rng = np.random.Generator(np.random.PCG64(seed))
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    cx, cy = (W / 2, H / 2) if center is None else center
    rgb = (rng.integers(0, 256, (H, W, 3)).astype(np.float32) * 0.5 + np.stack([xx / W, yy / H, (xx + yy) / (W + H)], -1) * 127).astype(np.uint8)
    a, b = extent[0] * H * 0.5, extent[1] * H * 0.5
    d = ((xx - cx) / a) ** 2 + ((yy - cy) / b) ** 2
    person = np.clip((1.05 - d) * 12, 0, 1)
    u, v = (xx - cx - 0.9 * a) * 0.8 + (yy - cy) * 0.6, -(xx - cx - 0.9 * a) * 0.6 + (yy - cy) * 0.8
    obj = np.clip((1.0 - np.maximum(np.abs(u) / (0.7 * a), np.abs(v) / (0.5 * a))) * 10, 0, 1)
    return rgb, (person * 255).astype(np.uint8), (obj * 255).astype(np.uint8)


def synthetic_recon_batch(frames: int = 96, size: int = 512, seed: int = 4, n_obj_points: int = 3000, obj_rings: int = 20, obj_segments: int = 20):
    """One batch of BASELINE config 4 (joint optimisation): everything ``recon_driver.fit_recon_batch`` reads, as host
    tensors.  Frames are temporally coherent where it matters for the optimiser (SMPL-T initialisation = a smooth random walk, object
    rotation = a slow drift, object and person masks = blobs that follow the projected object / body), the RGB content is noise (the
    network weights are random-init as well).

    Returns dict: images [T,8,S,S] fp32, crop_center [T,2], body_center [T,3], pose [T,156], betas [T,10], trans [T,3], body_kpts [T,25,3]
    (network-input pixels, confidence), obj_verts [V,3] / obj_faces [F,3] (closed template, ~800 faces as the BEHAVE templates),
    obj_points [n,3] (surface samples), obj_rot_init [T,3,3], occ_ratios [T]."""
    from .synth_smpl import synthetic_body_mesh, synthetic_motion
    rng = np.random.Generator(np.random.PCG64(seed))
    T, S = frames, size
    pose, betas, trans = synthetic_motion(T, seed=seed + 11)
    img = rng.random((T, 8, S, S), dtype=np.float32)
    img[:, 5:] = (img[:, 5:] > 0.5).astype(np.float32)
    crop = (np.array([[1024.0, 768.0]]) + np.cumsum(rng.standard_normal((T, 2)) * 3.0, 0)).astype(np.float32)
    body = (trans.numpy() + np.array([0.0, -0.25, 0.0])).astype(np.float32)             # body-25 joint 8 sits near the pelvis
    # object: an ellipsoid template 0.35 m beside the body, drifting slowly
    ov, of = synthetic_body_mesh(rings=obj_rings, segments=obj_segments, radii=(0.3, 0.25, 0.2))
    p = rng.standard_normal((n_obj_points, 3)); p /= np.linalg.norm(p, axis=1, keepdims=True)
    obj_points = (p * np.array([0.3, 0.25, 0.2])).astype(np.float32)
    obj_center = body + np.array([0.35, 0.0, 0.1], np.float32) + np.cumsum(rng.standard_normal((T, 3)) * 0.004, 0).astype(np.float32)
    ang = np.cumsum(rng.standard_normal((T, 3)) * 0.02, 0) + rng.standard_normal(3) * 0.3
    cx, sx, cy, sy, cz, sz = np.cos(ang[:, 0]), np.sin(ang[:, 0]), np.cos(ang[:, 1]), np.sin(ang[:, 1]), np.cos(ang[:, 2]), np.sin(ang[:, 2])
    Rm = np.zeros((T, 3, 3), np.float32)
    Rm[:, 0, 0], Rm[:, 0, 1], Rm[:, 0, 2] = cy * cz, -cy * sz, sy
    Rm[:, 1, 0], Rm[:, 1, 1], Rm[:, 1, 2] = sx * sy * cz + cx * sz, -sx * sy * sz + cx * cz, -sx * cy
    Rm[:, 2, 0], Rm[:, 2, 1], Rm[:, 2, 2] = -cx * sy * cz + sx * sz, cx * sy * sz + sx * cz, cx * cy
    # masks in network-input pixels: an ellipse where the object projects, a box around the body
    to_px = lambda c: ((600 + 979.7844 * c[:, 0] / c[:, 2] + 1018.952 - crop[:, 0]) * S / 1200, (600 + 979.840 * c[:, 1] / c[:, 2] + 779.486 - crop[:, 1]) * S / 1200)
    ox, oy = to_px(obj_center)
    bx, by = to_px(body)
    yy, xx = np.mgrid[0:S, 0:S].astype(np.float32)
    r_obj = 0.27 * 979.8 / obj_center[:, 2] * S / 1200
    obj_mask = (((xx[None] - ox[:, None, None]) / (1.15 * r_obj[:, None, None])) ** 2 + ((yy[None] - oy[:, None, None]) / (0.9 * r_obj[:, None, None])) ** 2) < 1
    hw, hh = 0.22 * 979.8 / body[:, 2] * S / 1200, 0.85 * 979.8 / body[:, 2] * S / 1200
    person_mask = (np.abs(xx[None] - bx[:, None, None]) < hw[:, None, None]) & (np.abs(yy[None] - by[:, None, None]) < hh[:, None, None])
    img[:, 3], img[:, 4] = person_mask.astype(np.float32), obj_mask.astype(np.float32)
    img[:, :3] *= np.maximum(img[:, 3:4], img[:, 4:5])
    # 2-D key points: a rigid joint cloud riding on the body centre, projected into the crop, + 2 px noise
    off = rng.standard_normal((1, 25, 3)) * np.array([0.25, 0.45, 0.12])
    J = body[:, None, :] + off
    kx = (600 + 979.7844 * J[..., 0] / J[..., 2] + 1018.952 - crop[:, 0:1]) * S / 1200 + rng.standard_normal((T, 25)) * 2
    ky = (600 + 979.840 * J[..., 1] / J[..., 2] + 779.486 - crop[:, 1:2]) * S / 1200 + rng.standard_normal((T, 25)) * 2
    kpts = np.stack([kx, ky, rng.uniform(0.3, 1.0, (T, 25))], -1).astype(np.float32)
    t = torch.from_numpy
    return {"images": t(img), "crop_center": t(crop), "body_center": t(body), "pose": pose, "betas": betas, "trans": trans, "body_kpts": t(kpts),
            "obj_verts": t(ov), "obj_faces": t(np.asarray(of, np.int64)), "obj_points": t(obj_points), "obj_rot_init": t(Rm),
            "occ_ratios": t(rng.uniform(0.4, 1.0, T).astype(np.float32))}
