"""Seeded inputs of the joint-optimisation parity tests (shared by make_golden.py, the oracle test and the GPU test)."""
import numpy as np
import torch

from fit_problem import load_assets
from oracle.smpl_ref import landmarks, smpl_forward
from vistracker_b200.synth import synthetic_frames
from vistracker_b200.synth_smpl import synthetic_motion, synthetic_smplh

B, N_OBJ = 4, 600


def make_problem(seed=31):
    a, reg = load_assets()
    rng = np.random.Generator(np.random.PCG64(seed))
    t = lambda x: torch.from_numpy(np.asarray(x, np.float32))
    model = synthetic_smplh(seed=3)
    pose, betas, trans = synthetic_motion(B, seed=seed)
    images, _, crop, _ = synthetic_frames(B, size=64, seed=seed, n_points=4, jitter=True)
    with torch.no_grad():
        verts = smpl_forward(model, pose, betas, trans)[0]
        J = landmarks(torch.stack([torch.as_tensor(reg[0]).long(), torch.as_tensor(reg[1]).long()]), torch.as_tensor(reg[2]), tuple(reg[3]), verts)
    body_center = J[:, 8].clone()
    labels = torch.from_numpy(a["part_labels"].astype(np.int64))
    pose_init = pose[:, 3:72] + t(rng.standard_normal((B, 69)) * 0.05)
    px = 600 + (979.7844 * J[..., 0] / J[..., 2] + 1018.952) - crop[:, 0:1]
    py = 600 + (979.840 * J[..., 1] / J[..., 2] + 779.486) - crop[:, 1:2]
    kp = torch.stack([px, py], -1) * 512 / 1200 + t(rng.standard_normal((B, 25, 2)) * 2)
    body_kpts = torch.cat([kp, t(rng.uniform(0.2, 1.0, (B, 25, 1)))], -1)
    p = rng.standard_normal((N_OBJ, 3)); p /= np.linalg.norm(p, axis=1, keepdims=True)
    objects = t(p * np.array([0.3, 0.25, 0.2]))[None].repeat(B, 1, 1)
    q = rng.standard_normal((B, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    Rm = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                   2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], 1).reshape(B, 3, 3)
    obj_R = t(Rm + rng.standard_normal((B, 3, 3)) * 0.01)
    obj_t = body_center + t(rng.standard_normal((B, 3)) * 0.1 + np.array([0.3, 0.0, 0.1]))
    d = dict(model=model, images=images, crop=crop, body_center=body_center, pose=pose, betas=betas, trans=trans, labels=labels,
             pose_init=pose_init, body_kpts=body_kpts, objects=objects, obj_R=obj_R, obj_t=obj_t, obj_s=torch.ones(B),
             occ=t(rng.uniform(0.3, 1.0, B)), noise=t(rng.random((B, 3, 3))), smpl_center=body_center.clone(),
             df_obj_h=t(rng.uniform(0.0, 0.2, (B, N_OBJ))), df_hum_o=t(rng.uniform(0.0, 0.25, (B, 6890))),
             parts_obj=torch.from_numpy(rng.integers(0, 14, (B, N_OBJ)).astype(np.int64)), assets=a, reg=reg)
    d["df_hum_o"][2] = 1.0           # a frame without human contacts is skipped (recon_fit_trivis_full.py:419-432)
    return d


def make_loop_extras(d, seed=57):
    """Extra inputs of the optimize_smpl_object loop golden (tests/golden/recon_obj_loop.npz): a low-poly closed template mesh (the
    reference's ``self.scan``), person / object masks whose object blob sits where the template projects, and the seeded U(0,1) draws that
    replace ``torch.rand`` inside ``decopose_axis`` in both runs."""
    from scipy.spatial import ConvexHull
    rng = np.random.Generator(np.random.PCG64(seed))
    t = lambda x: torch.from_numpy(np.asarray(x, np.float32))
    p = rng.standard_normal((28, 3)); p /= np.linalg.norm(p, axis=1, keepdims=True)
    tv = (p * np.array([0.3, 0.25, 0.2])).astype(np.float32)
    hull = ConvexHull(tv.astype(np.float64))
    tf = hull.simplices.astype(np.int64)
    # orient the faces outwards (ConvexHull does not promise a winding)
    c = tv.mean(0)
    for i, f in enumerate(tf):
        n = np.cross(tv[f[1]] - tv[f[0]], tv[f[2]] - tv[f[0]])
        if np.dot(n, tv[f[0]] - c) < 0:
            tf[i] = f[::-1]
    S = d["images"].shape[-1]
    crop, ot = d["crop"].numpy(), d["obj_t"].numpy()
    px = (600 + 979.7844 * ot[:, 0] / ot[:, 2] + 1018.952 - crop[:, 0]) * S / 1200
    py = (600 + 979.840 * ot[:, 1] / ot[:, 2] + 779.486 - crop[:, 1]) * S / 1200
    yy, xx = np.mgrid[0:S, 0:S]
    obj = np.stack([((xx - (x + 1.5)) ** 2 / 7.0 ** 2 + (yy - (y - 1.0)) ** 2 / 5.5 ** 2) < 1 for x, y in zip(px, py)]).astype(np.float32)
    person = np.zeros_like(obj); person[:, S // 4: 3 * S // 4, S // 2 - 6: S // 2 + 4] = 1.0
    images_sil = d["images"].clone()
    images_sil[:, 3], images_sil[:, 4] = t(person), t(obj)
    noise = t(rng.random((400, d["obj_R"].shape[0], 3, 3)))
    return dict(temp_v=tv, temp_f=tf, images_sil=images_sil, noise_seq=noise)
