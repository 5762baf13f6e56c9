"""Joint human-object optimisation on B200: the loss assembly and optimiser schedules of ``ReconFitterBehave.optimize_smpl`` /
``forward_smpl`` (recon/recon_fit_behave.py:393-513) and ``ReconFitterTriVisFull.optimize_smpl_object`` / ``forward_step``
(recon/recon_fit_trivis_full.py:124-457) over the hand-written kernels of this package:

    SMPL-H layer fwd/bwd (csrc/smpl.cu) · landmark regressors · SIF-Net fused query losses / query fwd (csrc/query_bwd_tc.cu, query_tc.cu) ·
    SO(3) projection, ragged Chamfer (csrc/geom.cu) · silhouette rasteriser fwd/bwd (csrc/raster.cu)

Two executions of the same loops:

* the default: every optimisation step is ONE CUDA-graph replay of a fixed kernel sequence (``recon_steps.SmplRefineStep`` /
  ``ObjectFitStep`` over csrc/recon.cu): loss terms, analytic gradients, masked Adam, loss history and the early-stop predicate all on the
  device, the host only walks the schedule;
* ``loop_mode = "eager"``: ``forward_smpl`` / ``forward_step`` below -- each operator a ``torch.autograd.Function`` around its kernels, the
  scalar glue and ``torch.optim.Adam`` in PyTorch -- written like the reference's methods; it is what the per-term goldens are checked
  against and the cross-check of the graph path (tests/test_gpu_recon_fit.py compares both with the reference's own loops).

File IO, data loading and mesh templates of the reference stay outside: everything is tensors.
"""
from __future__ import annotations

import struct
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from .geom import chamfer_distance_packed, chamfer_distance_ragged, decopose_axis, project_so3
from .recon_steps import OBJ_TERMS, RC_ESTOP, RC_LR0, RC_LR1, RC_PHASE, RC_SEED, RC_TEMP_K, RC_TOL, SMPL_TERMS, ObjectFitStep, SmplRefineStep
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
                 z_0: float = 2.2, net_in_size: int = 512, scan=None, loop_mode: str = "graph"):
        """scan: (vertices [V,3], faces [F,3]) of the centred object template (the reference's ``self.scan``), used to build the silhouette loss
        when the caller does not pass one.  loop_mode: 'graph' (one CUDA-graph replay per step) or 'eager' (PyTorch glue, the cross-check)."""
        self.model, self.priors, self.device = model, priors, model.device
        self.scan, self.loop_mode = scan, loop_mode
        self.last_hist, self.last_terms, self.last_stopped = None, None, False
        self.part_labels = part_labels.to(self.device).long()            # [6890]
        self.obj_scale, self.z_0, self.net_in_size = obj_scale, z_0, net_in_size
        self.collision_loss = False       # off unless the hostname matches two cluster nodes (recon_fit_base.py:106-108)

    # ------------------------------------------------------------------ schedules
    # recon_fit_trivis_full.py:124-153
    LOSS_WEIGHTS = {"beta": 1.0, "pose": 1e-5, "hand": 1e-5, "j2d": 0.3 ** 2, "object": 30.0 ** 2, "part": 0.05 ** 2, "contact": 30.0 ** 2,
                    "scale": 10.0 ** 2, "df_h": 10.0 ** 2, "smplz": 30 ** 2, "mask": 0.03 ** 2, "ocent": 0.0, "collide": 3 ** 2, "pinit": 5 ** 2,
                    "rot": 10.0 ** 2, "trans": 10.0 ** 2, "stemp": 100.0 ** 2, "otemp": 15.0 ** 2, "ovtemp": 50.0 ** 2}

    @classmethod
    def get_loss_weights(cls):
        return {k: (lambda cst, it, c=v: c * cst / (1 + it)) for k, v in cls.LOSS_WEIGHTS.items()}

    @staticmethod
    def sum_dict(loss_dict, weight_dict, it):
        return torch.stack([weight_dict[k](v, it) for k, v in loss_dict.items()]).sum()

    @staticmethod
    def get_opt_iters():
        return {"sil": 30, "object": 15}

    @staticmethod
    def smpl_phase_schedule(iter_for_betas, iter_for_pose, iter_for_kpts, max_iter):
        """The if / elif chain of recon_fit_behave.py:414-435 as a list of (phase, starts_new_optimizer) per outer iteration."""
        out, phase = [], None
        for it in range(iter_for_betas + iter_for_kpts + iter_for_pose + max_iter):
            new_opt = False
            if it < iter_for_betas:
                phase = "global"
            elif it == iter_for_betas:
                phase, new_opt = "smpl all pose", True
            elif it < iter_for_betas + iter_for_pose:
                pass
            elif it == iter_for_betas + iter_for_pose:
                phase = "kpts"
            out.append((phase, new_opt))
        return out

    @staticmethod
    def object_phase_schedule(it_obj, it_sil, joint_iter, max_iter):
        """The if / elif chain of recon_fit_trivis_full.py:329-348: (phase, starts_new_optimizer, decay) per outer iteration."""
        out, phase = [], None
        for it in range(joint_iter + it_obj + max_iter + it_sil):
            new_opt = False
            if it < it_obj:
                phase = "object only"
            elif it == it_obj and it != it_obj + it_sil:
                phase, new_opt = "sil", True
            elif it == it_obj + it_sil:
                phase, new_opt = "joint", True
            decay = 1 if phase == "object only" else it
            if phase == "sil":
                decay = it - it_obj + 1
            elif phase == "joint":
                decay = (it - it_obj + 1) / 3
            out.append((phase, new_opt, decay))
        return out

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

    def get_smpl_height(self, smpl: SMPLParams):
        """recon_fit_base.py:818-828."""
        with torch.no_grad():
            verts = smpl()[0]
        return verts[:, :, 1].amax(1) - verts[:, :, 1].amin(1)

    def optimize_smpl(self, smpl: SMPLParams, data_dict, iter_for_betas=10, iter_for_pose=10, iter_for_kpts=5, steps_per_iter=10,
                      max_iter=100, loop_mode: Optional[str] = None):
        """recon_fit_behave.py:393-465 (the tri-vis driver calls it with 1, 1, 1 -- recon_fit_triplane.py:66).  Returns ``(smpl, scale)`` like
        the reference: the container it was given, updated in place, and the body-height ratio after / before.

        In place is what the reference does, too: ``split_smpl`` -> ``SMPLPyTorchWrapperBatchSplitParams.from_smpl`` wraps VIEWS of
        ``smpl.pose.data / betas.data / trans.data`` in new Parameters (lib_smpl/wrapper_pytorch.py:206-226; ``.to()`` on the same device is a
        no-op), so Adam's in-place updates of the split parameters -- the eight 'other' betas included -- land in the caller's container, and
        ``copy_smpl_params`` re-copies a subset of the same storage (tests/golden/recon_loop.npz records ``alias = True`` from the reference's
        own run).  The per-step totals / terms are kept in ``self.last_hist`` / ``self.last_terms``, ``self.last_stopped``."""
        mode = self.loop_mode if loop_mode is None else loop_mode
        if mode == "eager":
            return self._optimize_smpl_eager(smpl, data_dict, iter_for_betas, iter_for_pose, iter_for_kpts, steps_per_iter, max_iter)
        n_it = iter_for_betas + iter_for_kpts + iter_for_pose + max_iter
        with torch.cuda.device(self.device):
            st = SmplRefineStep(self, smpl, data_dict, n_it * steps_per_iter)
            height_init = st.heights()
            self.last_stopped = st.run(self.LOSS_WEIGHTS, iter_for_betas, iter_for_pose, iter_for_kpts, steps_per_iter, max_iter)
            scale = st.heights() / height_init
            st.write_back(smpl)
            self.last_hist, self.last_terms = st._history()
            self.last_launches = st.steps_launched * st.LAUNCHES_PER_STEP
        return smpl, scale

    def _optimize_smpl_eager(self, smpl: SMPLParams, data_dict, iter_for_betas, iter_for_pose, iter_for_kpts, steps_per_iter, max_iter):
        """The same loop with PyTorch glue, autograd and torch.optim.Adam around the operator kernels (one host synchronisation per step for
        the early-stop test, which is evaluated on fp32 tensors exactly as recon_fit_behave.py:452 does)."""
        height_init = self.get_smpl_height(smpl)
        opt = torch.optim.Adam([smpl.top_betas, smpl.trans], lr=0.02)
        weight_dict = self.get_loss_weights()
        prev_loss, hist, terms = 300.0, [], []
        self.last_stopped = False

        def done():
            self.last_hist = np.asarray(hist, np.float64)
            self.last_terms = np.asarray([[float(ld[k]) if k in ld else np.nan for k in SMPL_TERMS] for ld in terms], np.float64)
            return smpl, self.get_smpl_height(smpl) / height_init
        for it, (phase, new_opt) in enumerate(self.smpl_phase_schedule(iter_for_betas, iter_for_pose, iter_for_kpts, max_iter)):
            if new_opt:
                opt = torch.optim.Adam([smpl.trans, smpl.global_pose, smpl.body_pose, smpl.top_betas, smpl.other_betas], 0.006, betas=(0.9, 0.999))
            for _ in range(steps_per_iter):
                opt.zero_grad()
                loss_dict = self.forward_smpl(smpl, data_dict, phase)
                decay = 1 if phase != "kpts" else it / 3
                loss = self.sum_dict(loss_dict, weight_dict, decay)
                loss.backward()
                opt.step()
                hist.append(float(loss)); terms.append({k: v.detach() for k, v in loss_dict.items()})
                if bool(abs(prev_loss - loss) / prev_loss < prev_loss * 0.001) and (it > 0.25 * max_iter + iter_for_betas + iter_for_pose):
                    self.last_stopped = True
                    return done()
                prev_loss = loss.detach()
        return done()

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
        """The (frame, body part) pairing of recon_fit_trivis_full.py:405-449 as packed index lists ``(h_idx, o_idx, h_off, o_off)``: rows of
        the flattened [B*Nh, 3] vertices / [B*No, 3] object points, grouped by (frame, part) in the reference's loop order (frame outer, part
        inner, point order kept), with int32 group offsets.  The contact masks are frozen after the first joint step (:242-253), so this runs
        ONCE per batch -- a handful of device-wide sorts / counts, no Python loop over frames and parts -- and every later step is one gather +
        one Chamfer launch (the reference re-runs its double loop, with a GPU->CPU sync per `in` test, on each of its ~1100 joint steps)."""
        mask_o, mask_h = df_obj_h < cont_thres, df_hum_o < cont_thres
        if part_o.dim() == 3:
            part_o = torch.argmax(part_o, 1)
        B, Nh, No = mask_h.shape[0], mask_h.shape[1], mask_o.shape[1]
        both = mask_h.any(1) & mask_o.any(1)                             # frames lacking either side are skipped (:419-432)
        labels_h = self.part_labels.to(mask_h.device)[None].expand(B, Nh)

        def side(mask, labels, n):
            idx = torch.nonzero((mask & both[:, None]).reshape(-1))[:, 0]               # ascending = frame-major, point order
            key = torch.div(idx, n, rounding_mode="floor") * SMPL_PARTS_NUM + labels.reshape(-1)[idx].long()
            srt = torch.sort(key, stable=True)
            return idx[srt.indices], srt.values
        hi, hk = side(mask_h, labels_h, Nh)
        oi, ok = side(mask_o, part_o, No)
        ch, co = torch.bincount(hk, minlength=B * SMPL_PARTS_NUM), torch.bincount(ok, minlength=B * SMPL_PARTS_NUM)
        valid = (ch > 0) & (co > 0)                                      # `if i not in label_h or i not in label_o: continue`
        if not bool(valid.any()):
            return None
        z = torch.zeros(1, dtype=torch.int64, device=hi.device)
        h_off = torch.cat([z, torch.cumsum(ch[valid], 0)]).to(torch.int32)
        o_off = torch.cat([z, torch.cumsum(co[valid], 0)]).to(torch.int32)
        return hi[valid[hk]], oi[valid[ok]], h_off, o_off

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

    def _silhouette(self, data_dict):
        """``SilLossROI(images[:, 3], images[:, 4], self.scan, crop_center, camera_params=..., crop_size=..., net_input_size=...)`` of
        recon_fit_trivis_full.py:289-294, unless the caller already put one into data_dict['silhouette']."""
        if data_dict.get("silhouette") is not None:
            return data_dict["silhouette"]
        if self.scan is None or "images" not in data_dict:
            raise ValueError("the 'sil' phase needs data_dict['silhouette'] or (fitter.scan, data_dict['images'])")
        from .render import SilLossROI
        images = data_dict["images"]
        sil = SilLossROI.from_masks(images[:, 3], images[:, 4], self.scan[0], self.scan[1], data_dict["query_dict"]["crop_center"],
                                    device=self.device, camera_params=data_dict.get("camera_params"), crop_size=data_dict.get("crop_size", 1200),
                                    net_input_size=data_dict.get("net_input_size", self.net_in_size))
        data_dict["silhouette"] = sil
        return sil

    def _first_joint_contacts(self, data_dict, object, smpl_verts):
        """recon_fit_trivis_full.py:242-253: the distance / part predictions that define the contact sets, evaluated once."""
        if "df_obj_h" not in data_dict:
            with torch.no_grad():
                self.model.query(object.detach(), **data_dict["query_dict"])
                preds = self.model.get_preds()
                df_obj_h, part_o = preds[0][:, 0, :].clone(), preds[2].clone()
                self.model.query(smpl_verts.detach(), **data_dict["query_dict"])
                data_dict["df_obj_h"], data_dict["df_hum_o"] = df_obj_h, self.model.get_preds()[0][:, 1, :].clone()
                data_dict["parts_obj"] = part_o
        if "contact_pairs" not in data_dict:
            data_dict["contact_pairs"] = self.contact_pairs(data_dict["df_hum_o"], data_dict["df_obj_h"], data_dict["parts_obj"])
        return data_dict["contact_pairs"]

    def optimize_smpl_object(self, model, data_dict, obj_iter=20, joint_iter=10, sil_iter=50, steps_per_iter=10, max_iter=100, noise_fn=None,
                             loop_mode: Optional[str] = None, seed: int = 1):
        """recon_fit_trivis_full.py:283-377, same signature (``model`` is the network the fitter already holds; ``data_dict['smpl']`` the body,
        ``obj_R / obj_t / obj_s``, ``objects``, ``occ_ratios``, ``query_dict`` and -- for the silhouette phase -- ``images`` or a ready
        ``silhouette``): 'object only' (Adam R lr .002, t lr .006) -> 'sil' (new Adam [R, t] .006) -> 'joint' (new Adam [t] .002) with the
        per-phase decay, the contact sets fixed on the first joint step and the joint-phase early stop.  ``max_iter`` is the reference's
        hard-wired 100.  ``noise_fn()`` -> [B,3,3] replays the U(0,1) draws of decopose_axis (parity runs); otherwise they are drawn on the
        device (Philox, ``seed``).  Returns ``(smpl, obj_R, obj_t)``; obj_R / obj_t are updated in place."""
        mode = self.loop_mode if loop_mode is None else loop_mode
        smpl = data_dict["smpl"]
        if model is not None and model is not self.model:
            raise ValueError("the fitter is bound to its own network (ReconFitterTriVisFull(model, ...))")
        sil = self._silhouette(data_dict)
        if mode == "eager":
            return self._optimize_smpl_object_eager(smpl, data_dict, joint_iter, steps_per_iter, max_iter, noise_fn)
        obj_R, obj_t = data_dict["obj_R"], data_dict["obj_t"]
        it_sil, it_obj = self.get_opt_iters()["sil"], self.get_opt_iters()["object"]
        sched = self.object_phase_schedule(it_obj, it_sil, joint_iter, max_iter)
        W = self.LOSS_WEIGHTS
        seed_word = struct.unpack("f", struct.pack("I", (int(seed) & 0x7FFFFF) | 1))[0]
        rows = []
        for it, (phase, _, decay) in enumerate(sched):
            row = [0.0] * 32
            on = {"otemp": True, "ovtemp": True, "mask": phase == "sil", "scale": True, "trans": phase == "sil", "object": phase != "sil",
                  "contact": phase == "joint"}
            for k, name in enumerate(OBJ_TERMS):
                row[k] = W[name] / (1 + decay) if on[name] else 0.0
            row[RC_LR0], row[RC_LR1] = (0.002, 0.006) if phase == "object only" else ((0.006, 0.006) if phase == "sil" else (0.0, 0.002))
            row[RC_PHASE] = float(ObjectFitStep.PHASES.index(phase))
            row[RC_TOL] = 0.0001
            row[RC_ESTOP] = 1.0 if (it > 0.25 * max_iter and phase == "joint") else 0.0
            row[RC_TEMP_K] = 10.0 if phase == "joint" else 1.0
            row[RC_SEED] = seed_word
            rows.append(row)
        with torch.cuda.device(self.device):
            with torch.no_grad():
                smpl_verts = smpl()[0].detach()
                data_dict["smpl_center"] = smpl.reg(smpl_verts)[:, 8]                  # compute_smpl_center_pred: body-25 joint 8
            st = ObjectFitStep(self, smpl_verts, data_dict, len(sched) * steps_per_iter, inject_noise=noise_fn is not None, seed=seed)
            st._upload_schedule(rows)
            st._start()
            stopped, contacts_ready = False, False
            for it, (phase, new_opt, _) in enumerate(sched):
                if new_opt:
                    st._new_optimizer()
                    if phase == "sil":
                        data_dict["rot_init"] = decopose_axis(obj_R.detach(), noise=None if noise_fn is None else noise_fn()).clone()
                        data_dict["trans_init"] = obj_t.detach().clone()
                        st.buf["t_init"].copy_(data_dict["trans_init"])
                st._set_row(it)
                pidx = ObjectFitStep.PHASES.index(phase)
                for _ in range(steps_per_iter):
                    if noise_fn is not None:
                        st.buf["noise"].copy_(noise_fn().reshape(-1, 9).to(self.device))
                    if phase == "joint" and not contacts_ready:
                        st.enqueue_pose()                                               # this step's R / object points: same draw as the replay below
                        st.set_contact_pairs(self._first_joint_contacts(data_dict, st.buf["object"], smpl_verts))
                        contacts_ready = True
                    st.step(pidx)
                    if rows[it][RC_ESTOP] != 0.0 and st._poll_stop():
                        stopped = True
                        break
                if stopped:
                    break
            torch.cuda.current_stream().synchronize()
            self.last_stopped = bool(st.ctrl[34].item() != 0.0)
            self.last_hist, self.last_terms = st._history()
            self.last_launches = st.steps_launched
        return smpl, obj_R, obj_t

    def _optimize_smpl_object_eager(self, smpl: SMPLParams, data_dict, joint_iter, steps_per_iter, max_iter, noise_fn):
        """The same loop with PyTorch glue, autograd and torch.optim.Adam around the operator kernels."""
        obj_R, obj_t, obj_s = data_dict["obj_R"], data_dict["obj_t"], data_dict["obj_s"]
        opt = torch.optim.Adam([{"params": obj_R, "lr": 0.002}, {"params": obj_t, "lr": 0.006}])
        weight_dict = self.get_loss_weights()
        it_sil, it_obj = self.get_opt_iters()["sil"], self.get_opt_iters()["object"]
        prev_loss, hist, terms = 300.0, [], []
        self.last_stopped = False
        data_dict["smpl_center"] = self.compute_smpl_center_pred(smpl)
        with torch.no_grad():
            data_dict["_smpl_verts_frozen"] = smpl()[0].detach()

        def done():
            data_dict.pop("_smpl_verts_frozen", None)
            self.last_hist = np.asarray(hist, np.float64)
            self.last_terms = np.asarray([[float(ld[k]) if k in ld else np.nan for k in OBJ_TERMS] for ld in terms], np.float64)
            return smpl, obj_R, obj_t
        for it, (phase, new_opt, decay) in enumerate(self.object_phase_schedule(it_obj, it_sil, joint_iter, max_iter)):
            if new_opt and phase == "sil":
                opt = torch.optim.Adam([obj_R, obj_t], lr=0.006)
                data_dict["rot_init"] = decopose_axis(obj_R, noise=None if noise_fn is None else noise_fn()).detach().clone()
                data_dict["trans_init"] = obj_t.detach().clone()
            elif new_opt:
                opt = torch.optim.Adam([obj_t], lr=0.002)
            for _ in range(steps_per_iter):
                opt.zero_grad()
                loss_dict = self.forward_step(smpl, data_dict, obj_R, obj_t, obj_s, phase, None if noise_fn is None else noise_fn())
                loss = self.sum_dict(loss_dict, weight_dict, decay)
                loss.backward()
                opt.step()
                hist.append(float(loss)); terms.append({k: v.detach() for k, v in loss_dict.items()})
                if bool(abs(prev_loss - loss) / prev_loss < prev_loss * 0.0001) and (it > 0.25 * max_iter) and phase == "joint":
                    self.last_stopped = True
                    return done()
                prev_loss = loss.detach()
        return done()

    def final_rotation(self, obj_R):
        """save_outputs stores the projection WITHOUT noise (recon_fit_base.py:303)."""
        return project_so3(obj_R.detach())
