"""Synthetic SMPL-H-shaped body model and motions.

The real ``SMPLH_male.pkl`` is licensed and not in the reference repository (README.md:42), so parity and throughput are
established on a random model of the real SHAPE: V = 6890 vertices, J = 52 joints (22 body + 2 x 15 hand), 10 shape
directions, 51 x 9 = 459 pose directions, the SMPL-H kinematic tree, skinning weights with 4 non-zeros per vertex
(as in SMPL) and a sparse-ish joint regressor with rows summing to one.  Buffer names follow
``SMPL_Layer.__init__`` (lib_smpl/smplpytorch/smplpytorch/pytorch/smpl_layer.py:50-71).
"""
from __future__ import annotations

import numpy as np
import torch

NUM_VERTS, NUM_JOINTS, NUM_BETAS = 6890, 52, 10
# kintree_table[0] of SMPL-H (22 body joints, then 15 left-hand and 15 right-hand joints hanging off the wrists 20 / 21)
SMPLH_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19,
                 20, 22, 23, 20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35,
                 21, 37, 38, 21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50]
assert len(SMPLH_PARENTS) == NUM_JOINTS


def synthetic_smplh(seed: int = 3, num_verts: int = NUM_VERTS):
    rng = np.random.Generator(np.random.PCG64(seed))
    V, J = num_verts, NUM_JOINTS
    f32 = np.float32
    v_template = (rng.standard_normal((1, V, 3)) * np.array([0.25, 0.45, 0.12])).astype(f32)
    shapedirs = (rng.standard_normal((V, 3, NUM_BETAS)) * 0.02).astype(f32)
    posedirs = (rng.standard_normal((V, 3, (J - 1) * 9)) * 0.004).astype(f32)
    jreg = np.zeros((J, V), f32)
    for j in range(J):
        idx = rng.choice(V, size=24, replace=False)
        w = rng.random(24).astype(f32)
        jreg[j, idx] = w / w.sum()
    weights = np.zeros((V, J), f32)
    for v in range(V):
        idx = rng.choice(J, size=4, replace=False)
        w = rng.random(4).astype(f32) + 0.05
        weights[v, idx] = w / w.sum()
    faces = rng.integers(0, V, size=(13776, 3)).astype(np.int64)
    return {"th_betas": torch.zeros(1, NUM_BETAS), "th_shapedirs": torch.from_numpy(shapedirs),
            "th_posedirs": torch.from_numpy(posedirs), "th_v_template": torch.from_numpy(v_template),
            "th_J_regressor": torch.from_numpy(jreg), "th_weights": torch.from_numpy(weights),
            "th_faces": torch.from_numpy(faces), "parents": list(SMPLH_PARENTS)}


# rest joints of the human-shaped model below, metres, y up, origin at the middle of the trunk: the 22 SMPL body joints in kintree order
_BODY_JOINTS = [(0, -0.05, 0), (0.07, -0.14, 0), (-0.07, -0.14, 0), (0, 0.07, 0), (0.10, -0.50, 0), (-0.10, -0.50, 0), (0, 0.20, 0),
                (0.09, -0.78, 0), (-0.09, -0.78, 0), (0, 0.27, 0), (0.10, -0.83, 0.08), (-0.10, -0.83, 0.08), (0, 0.48, 0),
                (0.06, 0.40, 0), (-0.06, 0.40, 0), (0, 0.60, 0), (0.17, 0.42, 0), (-0.17, 0.42, 0), (0.24, 0.18, 0), (-0.24, 0.18, 0),
                (0.26, -0.05, 0), (-0.26, -0.05, 0)]


def synthetic_smplh_surface(seed: int = 3, pelvis=(0.0, -0.25, 0.0)):
    """A HUMAN-SHAPED stand-in for ``SMPLH_male.pkl`` with the same buffers as ``synthetic_smplh``: the template is the closed 1.7 m surface of
    ``synthetic_body_mesh`` (6890 vertices ordered ring by ring, 13 776 coherent faces), the 52 rest joints sit where a person's do (22 body
    joints + 15 finger joints hanging off each wrist), each joint is regressed from the 24 template vertices nearest to it, every vertex is
    skinned to its four nearest joints (inverse-square-distance weights), the shape directions are smooth fields (one random 3 x 3 shear of
    the template per beta) and the pose directions a millimetre of noise.  What matters for throughput is what the real model has and
    ``synthetic_smplh`` (a Gaussian point cloud in random vertex order, right for skinning parity) has not: consecutive vertices are
    neighbours on a surface, so the 128 points of a query tile sample a patch of the feature maps instead of the whole body volume.
    ``pelvis``: where the root joint lies relative to the translation -- the offset the synthetic batches use for the body centre
    (vistracker_b200/synth.py: ``body = trans + (0, -0.25, 0)``), so the triplane is centred on the hips as in BEHAVE."""
    rng = np.random.Generator(np.random.PCG64(seed + 7000))
    f32 = np.float32
    J = NUM_JOINTS
    shift = np.asarray(pelvis, np.float64) - np.asarray(_BODY_JOINTS[0], np.float64)
    verts, faces = synthetic_body_mesh()
    v = verts.astype(np.float64) + shift
    V = v.shape[0]
    joints = np.zeros((J, 3))
    joints[:22] = np.asarray(_BODY_JOINTS, np.float64) + shift
    for side, wrist in ((0, 20), (1, 21)):                   # five fingers x three joints below each wrist
        for fgr in range(5):
            for k in range(3):
                joints[22 + side * 15 + fgr * 3 + k] = joints[wrist] + np.array([(0.012 if side == 0 else -0.012) * (fgr - 2), -0.03 * (k + 1) - 0.04, 0.01])
    jreg = np.zeros((J, V), f32)
    for j in range(J):
        idx = np.argsort(((v - joints[j]) ** 2).sum(1))[:24]
        w = rng.random(24).astype(f32) + 0.1
        jreg[j, idx] = w / w.sum()
    jpos = jreg.astype(np.float64) @ v                       # where the regressor puts the joints: the skinning neighbourhoods follow these
    d2 = ((v[:, None, :] - jpos[None]) ** 2).sum(-1)
    near = np.argsort(d2, 1)[:, :4]
    weights = np.zeros((V, J), f32)
    w = 1.0 / (np.take_along_axis(d2, near, 1) + 0.02 ** 2)
    np.put_along_axis(weights, near, (w / w.sum(1, keepdims=True)).astype(f32), 1)
    shear = rng.standard_normal((NUM_BETAS, 3, 3)) * 0.03
    shapedirs = (np.einsum("kab,vb->vak", shear, v - shift) + rng.standard_normal((V, 3, NUM_BETAS)) * 0.001).astype(f32)
    posedirs = (rng.standard_normal((V, 3, (J - 1) * 9)) * 0.001).astype(f32)
    return {"th_betas": torch.zeros(1, NUM_BETAS), "th_shapedirs": torch.from_numpy(shapedirs), "th_posedirs": torch.from_numpy(posedirs),
            "th_v_template": torch.from_numpy(v.astype(f32)[None]), "th_J_regressor": torch.from_numpy(jreg),
            "th_weights": torch.from_numpy(weights), "th_faces": torch.from_numpy(faces.astype(np.int64)), "parents": list(SMPLH_PARENTS)}


def synthetic_motion(frames: int, seed: int = 5, sigma: float = 0.02):
    """Smooth random walk in body pose: pose [T,156], betas [T,10] = (2.2, 0, ...)-like, trans [T,3]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pose = np.zeros((frames, 156), np.float32)
    pose[0, :66] = rng.standard_normal(66) * 0.3
    pose[0, 66:] = rng.standard_normal(90) * 0.1
    for t in range(1, frames):
        pose[t] = pose[t - 1]
        pose[t, :66] += rng.standard_normal(66) * sigma
    betas = np.tile((rng.standard_normal(NUM_BETAS) * 0.5).astype(np.float32), (frames, 1))
    trans = np.array([0.0, 0.0, 2.2], np.float32) + np.cumsum(rng.standard_normal((frames, 3)) * 0.01, 0).astype(np.float32)
    return torch.from_numpy(pose), torch.from_numpy(betas), torch.from_numpy(trans.astype(np.float32))


def synthetic_body_mesh(rings: int = 84, segments: int = 82, radii=(0.28, 0.85, 0.14)):
    """A closed, consistently oriented genus-0 surface with SMPL's vertex and face counts (rings * segments + 2 = 6890 vertices,
    2 * rings * segments = 13 776 faces): a latitude / longitude ellipsoid of body-like extent, vertices ordered ring by ring -- small,
    spatially coherent triangles like the real template's (``synthetic_smplh`` draws its faces at random: fine for skinning parity,
    meaningless for a rasteriser).  Returns (verts [V,3] float32, faces [F,3] int64)."""
    th = np.pi * (np.arange(rings) + 1) / (rings + 1)                       # polar angle of each ring, poles excluded
    ph = 2 * np.pi * np.arange(segments) / segments
    ring = np.stack([np.sin(th)[:, None] * np.cos(ph)[None], np.cos(th)[:, None] * np.ones_like(ph)[None], np.sin(th)[:, None] * np.sin(ph)[None]], -1)
    verts = np.concatenate([[[0.0, 1.0, 0.0]], ring.reshape(-1, 3), [[0.0, -1.0, 0.0]]], 0) * np.asarray(radii)
    vid = lambda r, s: 1 + r * segments + (s % segments)
    faces = []
    south = 1 + rings * segments
    for s in range(segments):
        faces.append((0, vid(0, s + 1), vid(0, s)))
        faces.append((south, vid(rings - 1, s), vid(rings - 1, s + 1)))
    for r in range(rings - 1):
        for s in range(segments):
            a, b, c, d = vid(r, s), vid(r, s + 1), vid(r + 1, s), vid(r + 1, s + 1)
            faces.append((a, b, c))
            faces.append((b, d, c))
    faces = np.asarray(faces, np.int64)
    assert verts.shape[0] == rings * segments + 2 and faces.shape[0] == 2 * rings * segments
    return verts.astype(np.float32), faces
