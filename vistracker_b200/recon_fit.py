"""Joint human-object optimisation on B200: the loss assembly and optimiser schedules of ``ReconFitterBehave.optimize_smpl`` /
``forward_smpl`` (recon/recon_fit_behave.py:393-513) and ``ReconFitterTriVisFull.optimize_smpl_object`` / ``forward_step``
(recon/recon_fit_trivis_full.py:124-457) over the hand-written kernels of this package:

    SMPL-H layer fwd/bwd (csrc/smpl.cu) · landmark regressors · SIF-Net fused query losses / query fwd (csrc/query_bwd_tc.cu, query_tc.cu) ·
    SO(3) projection, ragged Chamfer (csrc/geom.cu) · silhouette rasteriser fwd/bwd (csrc/raster.cu)

Each of those is a ``torch.autograd.Function`` around one or two kernel launches; the scalar glue between them (clamps, means,
the 14-way cross-entropy, the 63x63 prior products) and Adam are ordinary PyTorch device ops in this round -- see DESIGN.md
"what is not fused yet".  File IO, data loading and mesh templates of the reference stay outside: everything is tensors.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from .geom import chamfer_distance_packed, chamfer_distance_ragged, decopose_axis, project_so3
from .sifnet import CHORETriplaneVisibility
from .smpl import LandmarkRegressor, SMPL_Layer

SMPL_POSE_PRAMS_NUM, SMPL_PARTS_NUM = 72, 14            # lib_smpl/const.py


class SMPLParams:
    """``SMPLPyTorchWrapperBatchSplitParams`` (lib_smpl/wrapper_pytorch.py:94-226) without the nn.Module baggage: separately
    optimisable global_pose / body_pose / hand_pose / top_betas / other_betas / trans, ``forward()`` and ``get_landmarks()``."""

    def __init__(self, layer: SMPL_Layer, body25: LandmarkRegressor, pose, betas, trans):
        dev = layer.device
        P = lambda t: t.detach().clone().float().to(dev).requires_grad_(True)
        self.smpl, self.reg = layer, body25
        self.global_pose, self.body_pose, self.hand_pose = P(pose[:, :3]), P(pose[:, 3:66]), P(pose[:, 66:])
        self.top_betas, self.other_betas, self.trans = P(betas[:, :2]), P(betas[:, 2:]), P(trans)
        self.faces = layer.th_faces
        self.verts = self.jtr = None
        self.pose = torch.cat([self.global_pose, self.body_pose, self.hand_pose], 1)
        self.betas = torch.cat([self.top_betas, self.other_betas], 1)

    def forward(self):
        self.betas = torch.cat([self.top_betas, self.other_betas], 1)
        self.pose = torch.cat([self.global_pose, self.body_pose, self.hand_pose], 1)
        verts, jtr, tposed, naked = self.smpl(self.pose, th_betas=self.betas, th_trans=self.trans)
        self.verts, self.jtr = verts, jtr
        return verts, jtr, tposed, naked

    __call__ = forward

    def get_landmarks(self, use_cache=False):
        verts = self.verts if use_cache else self.forward()[0]
        return self.reg(verts), None, None          # only the body-25 set is consumed on this path

    @classmethod
    def from_smpl(cls, smpl: "SMPLParams") -> "SMPLParams":
        """``SMPLPyTorchWrapperBatchSplitParams.from_smpl`` (lib_smpl/wrapper_pytorch.py:206-226): fresh leaf parameters holding the values of
        another container (anything with ``pose [B,156]``, ``betas [B,10]``, ``trans [B,3]``, ``smpl``, ``reg``)."""
        return cls(smpl.smpl, smpl.reg, smpl.pose.data, smpl.betas.data, smpl.trans.data)

    @classmethod
    def get_smplh(cls, layer: SMPL_Layer, body25: LandmarkRegressor, poses, betas, trans, hand_mean=None) -> "SMPLParams":
        """``SMPLHGenerator.get_smplh`` (lib_smpl/smpl_generator.py:85-99): a container from a complete parameter set; 72-d SMPL poses are
        padded to the 156 SMPL-H values with the GRAB mean hand pose (``Priors.hand_mean`` / ``mean_hand_pose``)."""
        return cls(layer, body25, smplh_pose(poses, hand_mean), torch.as_tensor(np.asarray(betas), dtype=torch.float32), torch.as_tensor(np.asarray(trans), dtype=torch.float32))


SMPLH_POSE_PRAMS_NUM, SMPLH_HANDPOSE_START = 156, 66     # lib_smpl/const.py


def smplh_pose(poses, hand_mean=None) -> torch.Tensor:
    """The pose handling of ``SMPLHGenerator.get_smplh`` (lib_smpl/smpl_generator.py:88-96): [B,156] passes through, [B,72] gets the 90 hand
    values replaced by ``hand_mean`` (the body part keeps its first 66 values; SMPL's two hand joints are dropped)."""
    poses = torch.as_tensor(np.asarray(poses.detach().cpu() if torch.is_tensor(poses) else poses), dtype=torch.float32)
    if poses.shape[1] == SMPLH_POSE_PRAMS_NUM:
        return poses
    assert poses.shape[1] == SMPL_POSE_PRAMS_NUM, "using unknown source of smpl poses"
    if hand_mean is None:
        raise ValueError("72-d SMPL poses need the GRAB mean hand pose (Priors.hand_mean)")
    out = torch.zeros(poses.shape[0], SMPLH_POSE_PRAMS_NUM)
    out[:, :SMPL_POSE_PRAMS_NUM] = poses
    out[:, SMPLH_HANDPOSE_START:] = torch.as_tensor(np.asarray(hand_mean.detach().cpu() if torch.is_tensor(hand_mean) else hand_mean), dtype=torch.float32).reshape(-1)
    return out


def copy_smpl_params(split_smpl: SMPLParams, smpl: SMPLParams) -> SMPLParams:
    """``ReconFitterBase.copy_smpl_params`` (recon/recon_fit_base.py:808-816): the optimised global / body / hand pose, the two top betas and
    the translation written back into ``smpl`` (the other betas are left as they were)."""
    with torch.no_grad():
        smpl.global_pose.copy_(split_smpl.global_pose)
        smpl.body_pose.copy_(split_smpl.body_pose)
        smpl.hand_pose.copy_(split_smpl.hand_pose)
        smpl.top_betas.copy_(split_smpl.top_betas)
        smpl.trans.copy_(split_smpl.trans)
        smpl.pose = torch.cat([smpl.global_pose, smpl.body_pose, smpl.hand_pose], 1)
        smpl.betas = torch.cat([smpl.top_betas, smpl.other_betas], 1)
    return smpl


class Priors:
    """get_prior() / HandPrior(type='grab') (lib_smpl/th_smpl_prior.py:25-48, th_hand_prior.py:46-72) loaded ONCE (the reference
    re-reads three pickles from disk on every optimisation step)."""

    def __init__(self, arrays: Dict[str, np.ndarray], device):
        f = lambda a: torch.as_tensor(np.asarray(a, np.float32)).to(device)
        self.body_mean, self.body_prec = f(arrays["body_prior_mean"])[None], f(arrays["body_prior_precision"])
        self.hand_mean = torch.cat([f(arrays["lh_prior_mean"]), f(arrays["rh_prior_mean"])])[None]
        self.lh_prec, self.rh_prec = f(arrays["lh_prior_precision"])[None], f(arrays["rh_prior_precision"])[None]

    def pose(self, pose):
        t = torch.matmul(pose[:, 3:66] - self.body_mean, self.body_prec)
        return (t * t).sum(dim=1)

    def hand(self, pose):
        t = pose[:, 66:] - self.hand_mean
        l, r = torch.matmul(t[:, :45], self.lh_prec), torch.matmul(t[:, 45:], self.rh_prec)       # [1, B, 45] each
        t2 = torch.cat([l, r], dim=1)
        return (t2 * t2).sum(dim=1)                                                                # [1, 45] (sic)


class ReconFitterTriVisFull:
    def __init__(self, model: CHORETriplaneVisibility, priors: Priors, part_labels: torch.Tensor, obj_scale: float = 1.0,
                 z_0: float = 2.2, net_in_size: int = 512):
        self.model, self.priors, self.device = model, priors, model.device
        self.part_labels = part_labels.to(self.device).long()            # [6890]
        self.obj_scale, self.z_0, self.net_in_size = obj_scale, z_0, net_in_size
        self.collision_loss = False       # off unless the hostname matches two cluster nodes (recon_fit_base.py:106-108)

    # ------------------------------------------------------------------ schedules
    @staticmethod
    def get_loss_weights():
        """recon_fit_trivis_full.py:124-153."""
        w = {"beta": 1.0, "pose": 1e-5, "hand": 1e-5, "j2d": 0.3 ** 2, "object": 30.0 ** 2, "part": 0.05 ** 2, "contact": 30.0 ** 2,
             "scale": 10.0 ** 2, "df_h": 10.0 ** 2, "smplz": 30 ** 2, "mask": 0.03 ** 2, "ocent": 0.0, "collide": 3 ** 2, "pinit": 5 ** 2,
             "rot": 10.0 ** 2, "trans": 10.0 ** 2, "stemp": 100.0 ** 2, "otemp": 15.0 ** 2, "ovtemp": 50.0 ** 2}
        return {k: (lambda cst, it, c=v: c * cst / (1 + it)) for k, v in w.items()}

    @staticmethod
    def sum_dict(loss_dict, weight_dict, it):
        return torch.stack([weight_dict[k](v, it) for k, v in loss_dict.items()]).sum()

    @staticmethod
    def get_opt_iters():
        return {"sil": 30, "object": 15}

    # ------------------------------------------------------------------ SMPL refinement against the neural UDF
    def project_points(self, joints3d, crop_center):
        """recon_fit_base.py:787-796: body-25 joints into the 512x512 network-input image."""
        d = self.model.dims
        x, y, z = joints3d[:, :, 0:1], joints3d[:, :, 1:2], joints3d[:, :, 2:3]
        px = d.crop_size / 2 + (d.fx_px * x / z + d.cx_px) - crop_center[:, 0].unsqueeze(1).unsqueeze(1)
        py = d.crop_size / 2 + (d.fy_px * y / z + d.cy_px) - crop_center[:, 1].unsqueeze(1).unsqueeze(1)
        return torch.cat([px, py], -1) * self.net_in_size / d.crop_size

    def projection_loss(self, joints3d, joints2d, crop_center):
        proj = self.project_points(joints3d, crop_center)
        loss = F.mse_loss(proj[:, :, :2], joints2d[:, :, :2], reduction="none")
        return torch.mean(torch.sum(loss, axis=-1) * joints2d[:, :, 2])

    def forward_smpl(self, smpl: SMPLParams, data_dict, phase):
        """recon_fit_behave.py:467-513 with the tri-vis overrides (no smplz term; stemp from trivis_full.py:170-177)."""
        loss_dict = {}
        smpl_verts, _, _, _ = smpl()
        # torch.clamp(df_pred[:, 0:1], max=0.1).mean() and F.cross_entropy(parts_pred, labels, 'none').sum(-1).mean() of the reference,
        # as per-point terms + point gradients from ONE fused launch (vt_query_losses_tc) instead of query() + autograd
        vals_df, vals_ce = self.model.query_losses(smpl_verts, df_channel=0, clamp_max=0.1, part_labels=data_dict["part_labels"],
                                                   **data_dict["query_dict"])
        loss_dict["df_h"] = vals_df.mean()
        loss_dict["pose"] = torch.mean(self.priors.pose(smpl.pose))
        loss_dict["hand"] = torch.mean(self.priors.hand(smpl.pose))
        loss_dict["part"] = vals_ce.sum(-1).mean()
        # the reference runs a second (and in the 'kpts' phase a third) SMPL forward here through get_landmarks() -- smplz_loss is a
        # no-op in tri-vis and the parameters have not changed since smpl() above, so the cached vertices give the same landmarks
        loss_dict["pinit"] = torch.mean(torch.sum((smpl.pose[:, 3:SMPL_POSE_PRAMS_NUM] - data_dict["pose_init"]) ** 2, -1))
        if phase == "kpts":
            J, _, _ = smpl.get_landmarks(use_cache=True)
            loss_dict["j2d"] = self.projection_loss(J, data_dict["body_kpts"], data_dict["query_dict"]["crop_center"])
        if smpl_verts.shape[0] >= 4:
            v1, v2 = smpl_verts[1:-1] - smpl_verts[:-2], smpl_verts[2:] - smpl_verts[1:-1]
            loss_dict["stemp"] = F.mse_loss(v1, v2)
        return loss_dict

    def optimize_smpl(self, smpl: SMPLParams, data_dict, iter_for_betas=1, iter_for_pose=1, iter_for_kpts=1, steps_per_iter=10,
                      max_iter=100):
        """recon_fit_behave.py:393-465 (the tri-vis driver calls it with 1, 1, 1 -- recon_fit_triplane.py:66)."""
        opt = torch.optim.Adam([smpl.top_betas, smpl.trans], lr=0.02)
        weight_dict = self.get_loss_weights()
        prev_loss, phase, hist = 300.0, "global", []
        for it in range(iter_for_betas + iter_for_kpts + iter_for_pose + max_iter):
            if it < iter_for_betas:
                phase = "global"
            elif it == iter_for_betas:
                phase = "smpl all pose"
                opt = torch.optim.Adam([smpl.trans, smpl.global_pose, smpl.body_pose, smpl.top_betas, smpl.other_betas], 0.006, betas=(0.9, 0.999))
            if it == iter_for_betas + iter_for_pose:
                phase = "kpts"
            for _ in range(steps_per_iter):
                opt.zero_grad()
                loss_dict = self.forward_smpl(smpl, data_dict, phase)
                decay = 1 if phase != "kpts" else it / 3
                loss = self.sum_dict(loss_dict, weight_dict, decay)
                loss.backward()
                opt.step()
                lv = float(loss)
                hist.append(lv)
                if (abs(prev_loss - lv) / prev_loss < prev_loss * 0.001) and (it > 0.25 * max_iter + iter_for_betas + iter_for_pose):
                    return smpl, hist
                prev_loss = lv
        return smpl, hist

    # ------------------------------------------------------------------ object / joint optimisation
    @staticmethod
    def transform_obj_verts(verts, obj_R, obj_t, obj_s):
        return (torch.bmm(verts, obj_R) + obj_t.unsqueeze(1)) * obj_s.unsqueeze(1).unsqueeze(1)

    def temporal_loss_joint(self, obj_verts, loss_dict, phase):
        if obj_verts.shape[0] < 4:
            return
        weight = 10.0 if phase == "joint" else 1.0
        v1, v2 = obj_verts[1:-1] - obj_verts[:-2], obj_verts[2:] - obj_verts[1:-1]
        loss_dict["otemp"] = F.mse_loss(v1, v2) * weight
        loss_dict["ovtemp"] = F.mse_loss(obj_verts[1:], obj_verts[:-1]) * weight

    def contact_pairs(self, df_hum_o, df_obj_h, part_o, cont_thres=0.08):
        """The (frame, body part) pairing of recon_fit_trivis_full.py:405-449 as index lists.  The contact masks are frozen after
        the first joint step (:242-253), so the pairing is computed ONCE per batch and every later step is two gathers + one
        Chamfer launch (the reference re-runs the Python double loop with GPU->CPU syncs on each of its ~1100 joint steps)."""
        mask_o, mask_h = df_obj_h < cont_thres, df_hum_o < cont_thres
        if part_o.dim() == 3:
            part_o = torch.argmax(part_o, 1)
        B, Nh, No = mask_h.shape[0], mask_h.shape[1], mask_o.shape[1]
        dev = mask_h.device
        hi, oi, hoff, ooff = [], [], [0], [0]
        mask_h_c, mask_o_c, part_o_c, labels_c = mask_h.cpu(), mask_o.cpu(), part_o.cpu(), self.part_labels.cpu()
        for b in range(B):
            mh, mo = mask_h_c[b], mask_o_c[b]
            if int(mh.sum()) == 0 or int(mo.sum()) == 0:
                continue
            hv, ov = torch.nonzero(mh)[:, 0], torch.nonzero(mo)[:, 0]
            lh, lo = labels_c[hv], part_o_c[b][ov]
            for i in range(SMPL_PARTS_NUM):
                sh, so = hv[lh == i], ov[lo == i]
                if sh.numel() == 0 or so.numel() == 0:
                    continue
                hi.append(sh + b * Nh); oi.append(so + b * No)
                hoff.append(hoff[-1] + sh.numel()); ooff.append(ooff[-1] + so.numel())
        if not hi:
            return None
        return (torch.cat(hi).to(dev), torch.cat(oi).to(dev), torch.tensor(hoff, dtype=torch.int32, device=dev),
                torch.tensor(ooff, dtype=torch.int32, device=dev))

    def compute_contact_loss(self, df_hum_o, df_obj_h, object, smpl_verts, loss_dict, part_o, cont_thres=0.08, pairs="build"):
        """recon_fit_trivis_full.py:393-457: per frame and body part, pull human contact vertices and object contact points
        together with a bidirectional Chamfer distance over the ragged (frame, part) clouds -- one kernel launch."""
        if pairs == "build":
            pairs = self.contact_pairs(df_hum_o, df_obj_h, part_o, cont_thres)
        if pairs is None:
            return
        h_idx, o_idx, h_off, o_off = pairs
        hs = smpl_verts.reshape(-1, 3).index_select(0, h_idx)
        os_ = object.reshape(-1, 3).index_select(0, o_idx)
        loss_dict["contact"] = chamfer_distance_packed(hs, os_, h_off, o_off)

    def forward_step(self, smpl: SMPLParams, data_dict, obj_R, obj_t, obj_s, phase, noise: Optional[torch.Tensor] = None):
        """recon_fit_trivis_full.py:193-270.  ``noise`` replays the U(0,1) tensor of decopose_axis (parity runs)."""
        # the SMPL parameters are frozen throughout optimize_smpl_object (its optimisers hold obj_R / obj_t only): the driver computes
        # the vertices once and every step reuses them -- same values as the reference's per-step SMPL forward
        smpl_verts = data_dict["_smpl_verts_frozen"] if "_smpl_verts_frozen" in data_dict else smpl()[0]
        loss_dict = {}
        R = decopose_axis(obj_R, noise=noise)
        object = self.transform_obj_verts(data_dict["objects"], R, obj_t, obj_s)
        first_joint = phase == "joint" and "df_obj_h" not in data_dict
        if phase == "sil":
            # the reference still queries the network here (recon_fit_trivis_full.py:199-204), but no 'sil' loss term reads a prediction
            # (mask / scale / trans / temporal only): the launch is skipped
            df_pred = part_o = None
        elif first_joint:
            # first joint step: df_h and the part logits at the object points are needed for the contact masks
            self.model.query(object, **data_dict["query_dict"])
            preds = self.model.get_preds()
            df_pred, centers_pred_o, part_o = preds[0], preds[3], preds[2]
            vals_df_o = torch.clamp(df_pred[:, 1, :], max=0.8)
        else:
            # distance term + its gradient from the fused launch; the centre head (reported only, weight 0) rides along forward-only
            vals_df_o, _, extra = self.model.query_losses(object, df_channel=1, clamp_max=0.8, also=("centers",), **data_dict["query_dict"])
            centers_pred_o = extra["centers"]
            df_pred = part_o = None
        if phase != "sil":
            obj_center_pred = data_dict["smpl_center"] + torch.mean(centers_pred_o, -1)           # recon_fit_behave.py:370-380
        self.temporal_loss_joint(object, loss_dict, phase)
        if phase == "sil":
            sil = data_dict["silhouette"]
            per_frame, _ = sil(R, obj_t, obj_s, reduction="none")
            loss_dict["mask"] = (per_frame["mask"] * data_dict["occ_ratios"]).mean()
            loss_dict["scale"] = torch.mean((obj_s - self.obj_scale) ** 2)
            loss_dict["trans"] = torch.mean((obj_t - data_dict["trans_init"]) ** 2)
        else:
            loss_dict["object"] = (torch.mean(vals_df_o, -1) * data_dict["occ_ratios"]).mean()
            loss_dict["scale"] = torch.mean((obj_s - self.obj_scale) ** 2)
            # weight 0 in get_loss_weights ("no loss anymore"): the value is reported, but building its graph would drag the
            # centre head through the query backward for an identically-zero gradient
            with torch.no_grad():
                oc = torch.mean(object, 1)
                loss_dict["ocent"] = (F.mse_loss(oc, obj_center_pred, reduction="none").sum(-1) * data_dict["occ_ratios"]).mean()
            if phase == "joint":
                if "df_obj_h" not in data_dict:      # contact masks are computed once, on the first joint step (:242-253)
                    df_obj_h = df_pred[:, 0, :]
                    self.model.query(smpl_verts.detach(), **data_dict["query_dict"])
                    data_dict["df_obj_h"] = df_obj_h.detach()
                    data_dict["df_hum_o"] = self.model.get_preds()[0][:, 1, :].detach()
                    data_dict["parts_obj"] = part_o.detach()
                if "contact_pairs" not in data_dict:
                    data_dict["contact_pairs"] = self.contact_pairs(data_dict["df_hum_o"], data_dict["df_obj_h"], data_dict["parts_obj"])
                self.compute_contact_loss(data_dict["df_hum_o"], data_dict["df_obj_h"], object, smpl_verts, loss_dict,
                                          part_o=data_dict["parts_obj"], pairs=data_dict["contact_pairs"])
                if self.collision_loss:
                    raise NotImplementedError("the BVH collision term is dead on this path (hostname switch) and is not built")
        return loss_dict

    def compute_smpl_center_pred(self, smpl: SMPLParams):
        with torch.no_grad():
            J, _, _ = smpl.get_landmarks()
            return J[:, 8]

    def optimize_smpl_object(self, smpl: SMPLParams, data_dict, obj_iter=20, joint_iter=10, steps_per_iter=10, max_iter=100,
                             noise_fn=None):
        """recon_fit_trivis_full.py:283-377: 'object only' (Adam R lr .002, t lr .006) -> 'sil' (new Adam [R, t] .006) ->
        'joint' (new Adam [t] .002) with the per-phase decay schedule and the joint-phase early stop."""
        obj_R, obj_t, obj_s = data_dict["obj_R"], data_dict["obj_t"], data_dict["obj_s"]
        opt = torch.optim.Adam([{"params": obj_R, "lr": 0.002}, {"params": obj_t, "lr": 0.006}])
        weight_dict = self.get_loss_weights()
        it_sil, it_obj = self.get_opt_iters()["sil"], self.get_opt_iters()["object"]
        prev_loss, phase, hist = 300.0, "object only", []
        data_dict["smpl_center"] = self.compute_smpl_center_pred(smpl)
        with torch.no_grad():
            data_dict["_smpl_verts_frozen"] = smpl()[0].detach()
        for it in range(joint_iter + it_obj + max_iter + it_sil):
            if it < it_obj:
                phase = "object only"
            elif it == it_obj and it != it_obj + it_sil:
                phase = "sil"
                opt = torch.optim.Adam([obj_R, obj_t], lr=0.006)
                data_dict["rot_init"] = decopose_axis(obj_R, noise=None if noise_fn is None else noise_fn()).detach().clone()
                data_dict["trans_init"] = obj_t.detach().clone()
            elif it == it_obj + it_sil:
                phase = "joint"
                opt = torch.optim.Adam([obj_t], lr=0.002)
            for _ in range(steps_per_iter):
                opt.zero_grad()
                loss_dict = self.forward_step(smpl, data_dict, obj_R, obj_t, obj_s, phase, None if noise_fn is None else noise_fn())
                decay = 1 if phase == "object only" else it
                if phase == "sil":
                    decay = it - it_obj + 1
                elif phase == "joint":
                    decay = (it - it_obj + 1) / 3
                loss = self.sum_dict(loss_dict, weight_dict, decay)
                loss.backward()
                opt.step()
                lv = float(loss)
                hist.append(lv)
                if (abs(prev_loss - lv) / prev_loss < prev_loss * 0.0001) and (it > 0.25 * max_iter) and phase == "joint":
                    data_dict.pop("_smpl_verts_frozen", None)
                    return smpl, obj_R, obj_t, hist
                prev_loss = lv
        data_dict.pop("_smpl_verts_frozen", None)
        return smpl, obj_R, obj_t, hist

    def final_rotation(self, obj_R):
        """save_outputs stores the projection WITHOUT noise (recon_fit_base.py:303)."""
        return project_so3(obj_R.detach())
