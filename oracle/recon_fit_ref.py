"""CPU restatement of the joint human-object optimisation loops (TEST INFRASTRUCTURE: checker and CPU baseline only).

Follows, in plain PyTorch-CPU with autograd and torch.optim.Adam exactly as the reference composes them:

* ``ReconFitterBehave.optimize_smpl`` / ``forward_smpl``             recon/recon_fit_behave.py:393-513
  (``compute_df_h_loss`` / ``compute_prior_loss`` / ``compute_kpts_loss`` / ``projection_loss`` recon/recon_fit_base.py:625-647,767-802,
  ``temporal_loss_smpl`` recon/recon_fit_trivis_full.py:170-177, priors lib_smpl/th_smpl_prior.py:25-48, th_hand_prior.py:46-72)
* ``ReconFitterTriVisFull.optimize_smpl_object`` / ``forward_step``  recon/recon_fit_trivis_full.py:124-457
  (``decopose_axis`` / ``project_so3`` / ``transform_obj_verts`` recon/recon_fit_base.py:178-199,455-469, ``SilLossROI.forward``
  recon/obj_pose_roi.py:183-207, ``compute_contact_loss`` :393-457)

over the other restatements of this package (sifnet_ref: the network query, smpl_ref: the body model, geom_ref: SO(3) projection and
the ragged Chamfer distance, raster_ref: the silhouette rasteriser).

PINNED: both loops reproduce the reference's OWN loops run on the CPU -- per-step loss terms, totals, early stop, final parameters --
(tests/golden/recon_loop.npz, recon_obj_loop.npz; tests/test_oracle_recon_fit.py).  Inside them the rasteriser and the Chamfer operator are
third-party code restated from their published algorithms (parity unpinned, see raster_ref.py / geom_ref.py).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import raster_ref as RR
from .geom_ref import chamfer_ragged, project_so3, transform_obj_verts
from .sifnet_ref import sif_query
from .smpl_ref import landmarks, smpl_forward

# recon_fit_trivis_full.py:124-153
LOSS_W = {"beta": 1.0, "pose": 1e-5, "hand": 1e-5, "j2d": 0.3 ** 2, "object": 30.0 ** 2, "part": 0.05 ** 2, "contact": 30.0 ** 2, "scale": 10.0 ** 2,
          "df_h": 10.0 ** 2, "smplz": 30 ** 2, "mask": 0.03 ** 2, "ocent": 0.0, "collide": 3 ** 2, "pinit": 5 ** 2, "rot": 10.0 ** 2,
          "trans": 10.0 ** 2, "stemp": 100.0 ** 2, "otemp": 15.0 ** 2, "ovtemp": 50.0 ** 2}
SMPL_TERMS = ("df_h", "pose", "hand", "part", "pinit", "j2d", "stemp")
OBJ_TERMS = ("otemp", "ovtemp", "mask", "scale", "trans", "object", "ocent", "contact")
CAM = (979.7844, 979.840, 1018.952, 779.486, 1200.0)          # model/camera.py:26-40 in pixels of the 2048-wide image + crop size


def sum_dict(loss_dict, it):
    """ReconFitterBase.sum_dict with get_loss_weights: sum_k w_k * loss_k / (1 + it)."""
    return torch.stack([LOSS_W[k] * v / (1 + it) for k, v in loss_dict.items()]).sum()


class Problem:
    """Everything the loops read besides the optimised variables: network weights + cached maps, body model, regressor, priors, labels."""

    def __init__(self, sd, maps, model, reg, priors, part_labels, crop_center, body_center, net_in_size=512, cam=CAM):
        self.sd, self.maps, self.model, self.cam, self.net_in = sd, maps, model, cam, float(net_in_size)
        self.reg = (torch.stack([torch.as_tensor(reg[0]).long(), torch.as_tensor(reg[1]).long()]), torch.as_tensor(reg[2]), tuple(int(x) for x in reg[3]))
        f = lambda a: torch.as_tensor(np.asarray(a, np.float32))
        self.body_mean, self.body_prec = f(priors["body_prior_mean"]), f(priors["body_prior_precision"])
        self.hand_mean = torch.cat([f(priors["lh_prior_mean"]), f(priors["rh_prior_mean"])])
        self.lh_prec, self.rh_prec = f(priors["lh_prior_precision"]), f(priors["rh_prior_precision"])
        self.labels = torch.as_tensor(part_labels).long()
        self.cc, self.bc = crop_center.float(), body_center.float()

    def query(self, points):
        return sif_query(self.sd, self.maps, points, self.cc, self.bc, self.cam)

    def smpl(self, pose, betas, trans):
        return smpl_forward(self.model, pose, betas, trans, torch.zeros(pose.shape[0], self.model["th_v_template"].shape[1], 3))[0]

    def body25(self, verts):
        return landmarks(self.reg[0], self.reg[1], self.reg[2], verts)


# ------------------------------------------------------------------------------------------------ SMPL refinement
def forward_smpl(P: Problem, parts, pose_init, body_kpts, phase):
    """recon_fit_behave.py:467-513.  parts = dict(global_pose, body_pose, hand_pose, top_betas, other_betas, trans) of leaf tensors."""
    pose = torch.cat([parts["global_pose"], parts["body_pose"], parts["hand_pose"]], 1)
    betas = torch.cat([parts["top_betas"], parts["other_betas"]], 1)
    verts = P.smpl(pose, betas, parts["trans"])
    B = verts.shape[0]
    df, _, parts_pred, _, _ = P.query(verts)
    ld = {"df_h": torch.clamp(df[:, 0:1, :], max=0.1).mean()}
    t = torch.matmul(pose[:, 3:66] - P.body_mean[None], P.body_prec)
    ld["pose"] = torch.mean((t * t).sum(1))
    th = pose[:, 66:] - P.hand_mean[None]
    l, r = torch.matmul(th[:, :45], P.lh_prec[None]), torch.matmul(th[:, 45:], P.rh_prec[None])          # [1, B, 45] each (th_hand_prior.py:62-72)
    t2 = torch.cat([l, r], 1)
    ld["hand"] = torch.mean((t2 * t2).sum(1))
    ld["part"] = F.cross_entropy(parts_pred, P.labels[None].expand(B, -1), reduction="none").sum(-1).mean()
    ld["pinit"] = torch.mean(torch.sum((pose[:, 3:72] - pose_init) ** 2, -1))
    if phase == "kpts":
        J = P.body25(verts)
        fx, fy, cx, cy, crop = P.cam
        px = crop / 2 + (fx * J[:, :, 0:1] / J[:, :, 2:3] + cx) - P.cc[:, 0].unsqueeze(1).unsqueeze(1)
        py = crop / 2 + (fy * J[:, :, 1:2] / J[:, :, 2:3] + cy) - P.cc[:, 1].unsqueeze(1).unsqueeze(1)
        proj = torch.cat([px, py], -1) * P.net_in / crop
        ld["j2d"] = torch.mean(torch.sum(F.mse_loss(proj, body_kpts[:, :, :2], reduction="none"), -1) * body_kpts[:, :, 2])
    if B >= 4:
        ld["stemp"] = F.mse_loss(verts[1:-1] - verts[:-2], verts[2:] - verts[1:-1])
    return ld


def _height(P, parts):
    with torch.no_grad():
        v = P.smpl(torch.cat([parts["global_pose"], parts["body_pose"], parts["hand_pose"]], 1), torch.cat([parts["top_betas"], parts["other_betas"]], 1),
                   parts["trans"])
    return v[:, :, 1].max(1).values - v[:, :, 1].min(1).values


def optimize_smpl(P: Problem, pose, betas, trans, pose_init, body_kpts, iter_for_betas=10, iter_for_pose=10, iter_for_kpts=5, steps_per_iter=10,
                  max_iter=100, step_budget: Optional[int] = None):
    """recon_fit_behave.py:393-465.  Returns dict(pose, betas, trans, hist, terms, scale, stopped).  ``step_budget`` cuts the run after that
    many steps (bounded CPU-baseline samples)."""
    L = lambda t: t.detach().clone().float().requires_grad_(True)
    parts = {"global_pose": L(pose[:, :3]), "body_pose": L(pose[:, 3:66]), "hand_pose": L(pose[:, 66:]), "top_betas": L(betas[:, :2]),
             "other_betas": L(betas[:, 2:]), "trans": L(trans)}
    h0 = _height(P, parts)
    opt = torch.optim.Adam([parts["top_betas"], parts["trans"]], lr=0.02)
    prev_loss, hist, terms, stopped, phase = 300.0, [], [], False, None
    for it in range(iter_for_betas + iter_for_kpts + iter_for_pose + max_iter):
        if it < iter_for_betas:
            phase = "global"
        elif it == iter_for_betas:
            phase = "smpl all pose"
            opt = torch.optim.Adam([parts[k] for k in ("trans", "global_pose", "body_pose", "top_betas", "other_betas")], 0.006, betas=(0.9, 0.999))
        elif it < iter_for_betas + iter_for_pose:
            pass
        elif it == iter_for_betas + iter_for_pose:
            phase = "kpts"
        for _ in range(steps_per_iter):
            opt.zero_grad()
            ld = forward_smpl(P, parts, pose_init, body_kpts, phase)
            loss = sum_dict(ld, 1 if phase != "kpts" else it / 3)
            loss.backward()
            opt.step()
            hist.append(float(loss)); terms.append([float(ld[k]) if k in ld else np.nan for k in SMPL_TERMS])
            if bool(abs(prev_loss - loss) / prev_loss < prev_loss * 0.001) and (it > 0.25 * max_iter + iter_for_betas + iter_for_pose):
                stopped = True
                break
            prev_loss = loss.detach()
            if step_budget is not None and len(hist) >= step_budget:
                break
        if stopped or (step_budget is not None and len(hist) >= step_budget):
            break
    out = {k: v.detach() for k, v in parts.items()}
    return {"pose": torch.cat([out["global_pose"], out["body_pose"], out["hand_pose"]], 1), "betas": torch.cat([out["top_betas"], out["other_betas"]], 1),
            "trans": out["trans"], "hist": np.asarray(hist), "terms": np.asarray(terms), "scale": _height(P, parts) / h0, "stopped": stopped}


# ------------------------------------------------------------------------------------------------ silhouette loss
class _SilFn(torch.autograd.Function):
    """neural_renderer silhouettes + pseudo-gradient through raster_ref (numpy, float64)."""

    @staticmethod
    def forward(ctx, verts, faces, K, size):
        v, f = verts.detach().double().numpy(), np.asarray(faces)
        imgs, saved = [], []
        for b in range(v.shape[0]):
            K4 = (float(K[b, 0, 0]), float(K[b, 1, 1]), float(K[b, 0, 2]), float(K[b, 1, 2]))
            fv = RR.faces_of(RR.project(v[b], K4), f)
            idx, alpha, _ = RR.rasterize_fast(fv, size)
            imgs.append(alpha); saved.append((fv, idx, alpha, K4))
        ctx.saved, ctx.v, ctx.f, ctx.size = saved, v, f, size
        return torch.from_numpy(np.stack(imgs)).float()

    @staticmethod
    def backward(ctx, g):
        g = g.double().numpy()
        out = []
        for b, (fv, idx, alpha, K4) in enumerate(ctx.saved):
            out.append(RR.backward_verts(RR.backward_faces(fv, idx, alpha, g[b], ctx.size), ctx.v[b], ctx.f, K4))
        return torch.from_numpy(np.stack(out)).float(), None, None, None


class SilLoss:
    """SilLossROI.forward on ready-made ROI crops (keep mask, reference silhouette, ROI intrinsics): recon/obj_pose_roi.py:183-207."""

    def __init__(self, keep_mask, image_ref, K_roi, vertices, faces, rend_size=256):
        self.keep, self.ref, self.K = keep_mask.float(), image_ref.float(), K_roi.float()
        self.vertices, self.faces, self.size = torch.as_tensor(np.asarray(vertices), dtype=torch.float32), np.asarray(faces), rend_size

    def __call__(self, R, obj_t, obj_s):
        B = R.shape[0]
        verts = obj_s.view(-1, 1, 1) * (torch.bmm(self.vertices[None].expand(B, -1, -1), R) + obj_t.unsqueeze(1))
        image = self.keep * _SilFn.apply(verts, self.faces, self.K, self.size)
        return torch.sum((image - self.ref) ** 2, dim=(1, 2))


# ------------------------------------------------------------------------------------------------ object / joint optimisation
def contact_clouds(P: Problem, df_hum_o, df_obj_h, part_o, object, smpl_verts, thres=0.08):
    """The frame x part loop of compute_contact_loss (recon_fit_trivis_full.py:405-449) -> two lists of clouds."""
    mo_all, mh_all = df_obj_h < thres, df_hum_o < thres
    if part_o.dim() == 3:
        part_o = torch.argmax(part_o, 1)
    hs, os_ = [], []
    for hum, obj, mh, mo, po in zip(smpl_verts, object, mh_all, mo_all, part_o):
        if int(mh.sum()) == 0 or int(mo.sum()) == 0:
            continue
        obj_v, label_o, hum_v, label_h = obj[mo], po[mo], hum[mh], P.labels[mh]
        for i in range(14):
            if i not in label_h or i not in label_o:
                continue
            hs.append(hum_v[torch.where(label_h == i)[0]]); os_.append(obj_v[torch.where(label_o == i)[0]])
    return hs, os_


def forward_step(P: Problem, smpl_verts, state: Dict, obj_R, obj_t, obj_s, phase, noise):
    """recon_fit_trivis_full.py:193-270.  state: objects [B,N,3], occ [B], smpl_center [B,3], sil (SilLoss), trans_init, obj_scale and the
    contact sets once they exist."""
    ld = {}
    R = project_so3(obj_R + 1e-4 * noise)
    object = transform_obj_verts(state["objects"], R, obj_t, obj_s)
    df, _, part_o, centers, _ = P.query(object)
    occ = state["occ"]
    if object.shape[0] >= 4:
        w = 10.0 if phase == "joint" else 1.0
        ld["otemp"] = F.mse_loss(object[1:-1] - object[:-2], object[2:] - object[1:-1]) * w
        ld["ovtemp"] = F.mse_loss(object[1:], object[:-1]) * w
    if phase == "sil":
        ld["mask"] = (state["sil"](R, obj_t, obj_s) * occ).mean()
        ld["scale"] = torch.mean((obj_s - state["obj_scale"]) ** 2)
        ld["trans"] = torch.mean((obj_t - state["trans_init"]) ** 2)
    else:
        ld["object"] = (torch.mean(torch.clamp(df[:, 1, :], max=0.8), -1) * occ).mean()
        ld["scale"] = torch.mean((obj_s - state["obj_scale"]) ** 2)
        center_pred = state["smpl_center"] + torch.mean(centers, -1)
        ld["ocent"] = (F.mse_loss(torch.mean(object, 1), center_pred, reduction="none").sum(-1) * occ).mean()
        if phase == "joint":
            if "df_obj_h" not in state:
                with torch.no_grad():
                    df_h2 = P.query(smpl_verts)[0]
                state["df_obj_h"], state["df_hum_o"], state["parts_obj"] = df[:, 0, :].detach(), df_h2[:, 1, :].detach(), part_o.detach()
            hs, os_ = contact_clouds(P, state["df_hum_o"], state["df_obj_h"], state["parts_obj"], object, smpl_verts)
            if hs:
                ld["contact"] = chamfer_ragged(hs, os_)
    return ld


def optimize_smpl_object(P: Problem, pose, betas, trans, obj_R, obj_t, obj_s, objects, occ, sil: SilLoss, noise_fn: Callable, it_obj=15, it_sil=30,
                         joint_iter=10, steps_per_iter=10, max_iter=100, obj_scale=1.0, step_budget: Optional[Dict[str, int]] = None):
    """recon_fit_trivis_full.py:283-377.  noise_fn() -> [B,3,3] U(0,1) draws (one per decopose_axis call, in the reference's order).
    ``step_budget`` = {'object only': n, 'sil': n, 'joint': n} runs at most n steps of each phase and then jumps to the next one (bounded
    CPU-baseline samples; the result is then not the full optimisation).  Returns dict(obj_R, obj_t, hist, terms, phases, stopped, state)."""
    with torch.no_grad():
        smpl_verts = P.smpl(pose.float(), betas.float(), trans.float())
        smpl_center = P.body25(smpl_verts)[:, 8]
    obj_R, obj_t = obj_R.detach().clone().float().requires_grad_(True), obj_t.detach().clone().float().requires_grad_(True)
    obj_s = obj_s.float()
    state = {"objects": objects.float(), "occ": occ.float(), "smpl_center": smpl_center, "sil": sil, "obj_scale": obj_scale}
    opt = torch.optim.Adam([{"params": obj_R, "lr": 0.002}, {"params": obj_t, "lr": 0.006}])
    prev_loss, hist, terms, phases, stopped, phase = 300.0, [], [], [], False, None
    used = {"object only": 0, "sil": 0, "joint": 0}
    for it in range(joint_iter + it_obj + max_iter + it_sil):
        if it < it_obj:
            phase = "object only"
        elif it == it_obj and it != it_obj + it_sil:
            phase = "sil"
            opt = torch.optim.Adam([obj_R, obj_t], lr=0.006)
            state["rot_init"] = project_so3(obj_R + 1e-4 * noise_fn()).detach().clone()
            state["trans_init"] = obj_t.detach().clone()
        elif it == it_obj + it_sil:
            phase = "joint"
            opt = torch.optim.Adam([obj_t], lr=0.002)
        for _ in range(steps_per_iter):
            if step_budget is not None and used[phase] >= step_budget.get(phase, 0):
                break
            used[phase] += 1
            opt.zero_grad()
            ld = forward_step(P, smpl_verts, state, obj_R, obj_t, obj_s, phase, noise_fn())
            decay = 1 if phase == "object only" else it
            if phase == "sil":
                decay = it - it_obj + 1
            elif phase == "joint":
                decay = (it - it_obj + 1) / 3
            loss = sum_dict(ld, decay)
            loss.backward()
            opt.step()
            hist.append(float(loss)); terms.append([float(ld[k]) if k in ld else np.nan for k in OBJ_TERMS]); phases.append(phase)
            if bool(abs(prev_loss - loss) / prev_loss < prev_loss * 0.0001) and (it > 0.25 * max_iter) and phase == "joint":
                stopped = True
                break
            prev_loss = loss.detach()
        if stopped:
            break
    return {"obj_R": obj_R.detach(), "obj_t": obj_t.detach(), "rot_final": project_so3(obj_R.detach()), "hist": np.asarray(hist), "terms": np.asarray(terms),
            "phases": phases, "stopped": stopped, "state": state, "smpl_verts": smpl_verts}
