"""One optimisation step of the joint human-object fit as ONE CUDA-graph replay (csrc/recon.cu + the operator kernels):

* ``SmplRefineStep``  -- a step of ``ReconFitterBehave.optimize_smpl`` (recon/recon_fit_behave.py:437-458): SMPL-H forward -> fused SIF-Net
  query losses at the 6890 vertices -> loss terms and analytic gradients -> SMPL-H backward -> masked Adam -> history row + early-stop test;
* ``ObjectFitStep``   -- a step of ``ReconFitterTriVisFull.optimize_smpl_object`` (recon/recon_fit_trivis_full.py:350-374) in its three
  phases ('object only', 'sil', 'joint'): decopose_axis noise -> SO(3) projection -> rigid transform -> query losses / silhouette render /
  ragged Chamfer -> gradients back to obj_R, obj_t -> Adam -> history row + early-stop test.

All state lives in static device buffers; the schedule (per-term weights / (1 + decay), learning rates, phase, early-stop window) is a
table of rows uploaded once, one row copied device-to-device per outer iteration.  The host never waits for a step: inside the early-stop
window it reads the stop flag of the step before last through pinned memory, and the device-side guard (csrc/recon.cu) makes the one or
two replays queued past the stop no-ops, so the result is exactly the reference's `return` at that step.
"""
from __future__ import annotations

import collections
import ctypes
import os
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib

P, S = _lib.ptr, _lib.stream_ptr

# csrc/recon.cu control-block words
RC_LR0, RC_LR1, RC_PHASE, RC_TOL, RC_ESTOP, RC_TEMP_K, RC_SEED = 16, 17, 18, 19, 20, 21, 22
RC_STEP, RC_HIST, RC_STOP, RC_PREV, RC_DRAW = 32, 33, 34, 35, 36
SMPL_TERMS = ("df_h", "pose", "hand", "part", "pinit", "j2d", "stemp")
OBJ_TERMS = ("otemp", "ovtemp", "mask", "scale", "trans", "object", "contact")


def _zero(t: torch.Tensor):
    _lib.call("vt_zero", P(t), t.numel() * t.element_size(), S())


def capture_graph(enqueue) -> torch.cuda.CUDAGraph:
    """Capture ``enqueue()`` on a side stream with ``CUDAGraph.capture_begin / capture_end`` directly.  The ``torch.cuda.graph`` context
    manager also runs ``gc.collect()`` and ``torch.cuda.empty_cache()`` on entry, which hands every cached block back to the driver:
    measured 0.11 s per capture on a B200 holding the feature maps of a 96-frame batch (four captures per batch: 12 % of the batch time),
    plus the re-allocation cost afterwards."""
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        g.capture_begin()
        try:
            enqueue()
        finally:
            g.capture_end()
    torch.cuda.current_stream().wait_stream(side)
    return g


class _GraphLoop:
    """Shared plumbing: control block, history, schedule table, graph capture with state save / restore, asynchronous early-stop polling."""

    TERMS: tuple = ()

    def _init_ctrl(self, device, max_hist: int):
        lib = _lib.load()
        self.device = device
        self.ld = lib.vt_recon_hist_ld()
        self.ctrl = torch.zeros(lib.vt_recon_ctrl_words(), dtype=torch.float32, device=device)
        self.acc = torch.zeros(8, dtype=torch.float64, device=device)
        self.hist = torch.zeros(max_hist, self.ld, dtype=torch.float64, device=device)
        self.max_hist = max_hist
        self._flag_host = torch.zeros(4, dtype=torch.float32).pin_memory()
        self._pending = collections.deque()
        self._slot = 0
        self.graphs: Dict[int, torch.cuda.CUDAGraph] = {}
        self.steps_launched = 0

    def _upload_schedule(self, rows):
        """rows: list of 32-word lists -> device table; ``_set_row(i)`` copies one into ctrl[0:32] on the stream."""
        self.sched = torch.tensor(rows, dtype=torch.float32).to(self.device)

    def _set_row(self, i: int):
        self.ctrl[:32].copy_(self.sched[i])

    def _new_optimizer(self):
        """A NEW torch.optim.Adam in the reference: moments and step count start from zero."""
        self.m.zero_(); self.v.zero_(); self.ctrl[RC_STEP:RC_STEP + 1].zero_()

    def _start(self, prev_loss: float = 300.0):
        self.ctrl.zero_(); self.acc.zero_(); self.hist.zero_()
        self.ctrl[RC_PREV:RC_PREV + 1].fill_(prev_loss)
        self._pending.clear()
        self.steps_launched = 0

    def _mutable_state(self):
        raise NotImplementedError

    def _capture(self, key: int, enqueue):
        """Warm-up launch outside capture (lazy module loading, function attributes), capture, restore the state the warm-up changed."""
        keep = [t.clone() for t in self._mutable_state()]
        enqueue()
        torch.cuda.current_stream().synchronize()
        g = capture_graph(enqueue)
        with torch.no_grad():
            for t, k in zip(self._mutable_state(), keep):
                t.copy_(k)
        self.graphs[key] = g
        return g

    def _run(self, key: int, enqueue, use_graph: bool):
        if not use_graph:
            enqueue()
        else:
            g = self.graphs.get(key)
            if g is None:
                g = self._capture(key, enqueue)
            g.replay()
        self.steps_launched += 1

    def _poll_stop(self, lag: int = 2) -> bool:
        """Queue an asynchronous read-back of the stop flag after the step just launched; look at the flag of the step ``lag`` launches ago
        (its copy has long completed while the GPU works on the newer steps, so the wait does not drain the queue)."""
        slot = self._slot
        self._slot = (self._slot + 1) % self._flag_host.numel()
        self._flag_host[slot:slot + 1].copy_(self.ctrl[RC_STOP:RC_STOP + 1], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._pending.append((ev, slot))
        while len(self._pending) > lag:
            ev, s = self._pending.popleft()
            ev.synchronize()
            if float(self._flag_host[s]) != 0.0:
                return True
        return False

    def _history(self):
        """(totals [n], terms [n, len(TERMS)]) of the steps actually taken (rows written before the stop)."""
        n = int(self.ctrl[RC_HIST].item())
        h = self.hist[:min(n, self.max_hist)].cpu().numpy()
        return h[:, 14].copy(), h[:, :len(self.TERMS)].copy()


class SmplRefineStep(_GraphLoop):
    TERMS = SMPL_TERMS

    def __init__(self, fitter, smpl, data_dict, max_hist: int):
        net, layer, reg = fitter.model, smpl.smpl, smpl.reg
        dev = layer.device
        self.fitter, self.net, self.layer, self.reg = fitter, net, layer, reg
        self._init_ctrl(dev, max_hist)
        B = smpl.trans.shape[0]
        m = layer
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        self.B = B
        b = self.buf = {}
        b["pose"], b["betas"], b["trans"] = f(B, 156), f(B, 10), f(B, 3)
        b["coef"], b["R"], b["J"], b["G"], b["A"] = f(B, m.kdp), f(B, m.J, 9), f(B, m.J, 3), f(B, m.J, 12), f(B, m.J, 12)
        b["naked"], b["verts"], b["jtr"] = f(B, m.V, 3), f(B, m.V, 3), f(B, m.J, 3)
        b["J25"], b["gJ25"], b["g_verts"] = f(B, reg.L, 3), f(B, reg.L, 3), f(B, m.V, 3)
        b["g_vposed"], b["gA"], b["g_coef"], b["g_ts"] = f(B, m.nv3p), f(B, m.J, 12), f(B, m.kdp), f(B, 3)
        b["g_pose"], b["g_pose_d"], b["g_betas"], b["g_trans"] = f(B, 156), f(B, 156), f(B, 10), f(B, 3)
        b["vals_df"], b["g_df"], b["vals_ce"], b["g_ce"] = f(B, m.V), f(B, m.V, 3), f(B, m.V), f(B, m.V, 3)
        self.m, self.v = torch.zeros(B, 169, device=dev), torch.zeros(B, 169, device=dev)
        to = dict(device=dev, dtype=torch.float32)
        qd = data_dict["query_dict"]
        self.cc, self.bc = qd["crop_center"].to(**to).contiguous(), qd["body_center"].to(**to).contiguous()
        self.labels = data_dict["part_labels"].to(dev, torch.int64).contiguous()
        if tuple(self.labels.shape) != (B, m.V):
            raise ValueError(f"part_labels must be [B, V] = {(B, m.V)}, got {tuple(self.labels.shape)}")
        self.kpts = data_dict["body_kpts"].to(**to).contiguous()
        self.pose_init = data_dict["pose_init"].to(**to).contiguous()
        if tuple(self.pose_init.shape) != (B, 69):
            raise ValueError(f"pose_init must be [B, 69] (pose[:, 3:72]), got {tuple(self.pose_init.shape)}")
        pr = fitter.priors
        self._pri = [pr.body_mean.reshape(-1).contiguous(), pr.body_prec.contiguous(), pr.hand_mean.reshape(-1)[:45].contiguous(),
                     pr.lh_prec.reshape(45, 45).contiguous(), pr.hand_mean.reshape(-1)[45:].contiguous(), pr.rh_prec.reshape(45, 45).contiguous()]
        d = net.dims
        self._cam6 = (ctypes.c_float * 6)(d.fx_px, d.fy_px, d.cx_px, d.cy_px, d.crop_size, fitter.net_in_size)
        V = m.V
        self._div = (ctypes.c_double * 8)(B * V, B, 45.0, B, B, B * reg.L, max(B - 2, 1) * 3.0 * V, 1.0)
        self._maps = net._maps                      # the graph holds these pointers: keep the tensors alive as long as the graph
        self.merge_heads = os.environ.get("VT_QUERY_MERGE", "1") != "0"
        with torch.no_grad():
            b["pose"].copy_(torch.cat([smpl.global_pose, smpl.body_pose, smpl.hand_pose], 1))
            b["betas"].copy_(torch.cat([smpl.top_betas, smpl.other_betas], 1))
            b["trans"].copy_(smpl.trans)

    def _mutable_state(self):
        b = self.buf
        return [b["pose"], b["betas"], b["trans"], self.m, self.v, self.ctrl, self.acc, self.hist]

    def enqueue_forward(self):
        """SMPL-H forward only (heights before / after the fit)."""
        b, m = self.buf, self.layer
        _lib.call("vt_smpl_fwd", ctypes.byref(m.struct), P(b["pose"]), P(b["betas"]), P(b["trans"]), None, 1.0, self.B, P(b["coef"]), P(b["R"]),
                  P(b["J"]), P(b["G"]), P(b["A"]), P(b["naked"]), P(b["naked"]), P(b["verts"]), P(b["jtr"]), S())

    def _enqueue_step(self):
        b, m, B, reg = self.buf, self.layer, self.B, self.reg
        ms = ctypes.byref(m.struct)
        ctrl, acc = P(self.ctrl), P(self.acc)
        self.enqueue_forward()
        if self.merge_heads:
            # df_h and part heads merged in the kernel: the schedule's weights (ctrl words 0 and 3) times 1 / (B V) and 1 / B are folded into the
            # cotangents, one backward gather serves both terms, g_df holds w_dfh d df_h + w_part d part
            c0 = self.ctrl.data_ptr()
            self.net.enqueue_query_losses_merged(b["verts"], self.cc, self.bc, 0, 0.1, self.labels, c0, 1.0 / (B * m.V), c0 + 4 * 3, 1.0 / B,
                                                 b["vals_df"], b["vals_ce"], b["g_df"], maps=self._maps)
            _lib.call("vt_recon_point_terms", P(b["verts"]), B, m.V, 6, -1, 6, 0, 0, P(b["vals_df"]), P(b["g_df"]), None, -2, 0, 1.0,
                      P(b["vals_ce"]), None, -1, 3, 1.0, ctrl, P(b["g_verts"]), acc, S())
        else:
            self.net.enqueue_query_losses(b["verts"], self.cc, self.bc, 0, 0.1, self.labels, b["vals_df"], b["g_df"], b["vals_ce"], b["g_ce"],
                                          maps=self._maps)
            # stemp + df_h + part -> g_verts
            _lib.call("vt_recon_point_terms", P(b["verts"]), B, m.V, 6, -1, 6, 0, 0, P(b["vals_df"]), P(b["g_df"]), None, 0, 0, float(B * m.V),
                      P(b["vals_ce"]), P(b["g_ce"]), 3, 3, float(B), ctrl, P(b["g_verts"]), acc, S())
        _lib.call("vt_landmarks_fwd", P(b["verts"]), B, m.V, P(reg.rowptr), P(reg.col), P(reg.val), reg.L, P(b["J25"]), S())
        _lib.call("vt_recon_kpts", P(b["J25"]), P(self.kpts), P(self.cc), B, reg.L, self._cam6, ctrl, P(b["gJ25"]), acc, S())
        _lib.call("vt_landmarks_bwd", P(b["gJ25"]), B, m.V, P(reg.rowptr), P(reg.col), P(reg.val), reg.L, P(b["g_verts"]), S())
        _lib.call("vt_recon_pose_terms", P(b["pose"]), P(self.pose_init), B, *(P(t) for t in self._pri), ctrl, P(b["g_pose_d"]), acc, S())
        _lib.call("vt_smpl_bwd", ms, P(b["pose"]), P(b["R"]), P(b["J"]), P(b["G"]), P(b["A"]), P(b["naked"]), P(b["g_verts"]), None, 1.0, B,
                  P(b["g_vposed"]), P(b["gA"]), P(b["g_coef"]), P(b["g_ts"]), P(b["g_pose"]), P(b["g_betas"]), P(b["g_trans"]), S())
        _lib.call("vt_recon_adam_smpl", P(b["pose"]), P(b["betas"]), P(b["trans"]), P(b["g_pose"]), P(b["g_pose_d"]), P(b["g_betas"]),
                  P(b["g_trans"]), P(self.m), P(self.v), B, ctrl, S())
        _lib.call("vt_recon_end_step", acc, len(SMPL_TERMS), self._div, -1, None, ctrl, P(self.hist), self.max_hist, S())

    LAUNCHES_PER_STEP = 16          # kernels of one step (SMPL-H fwd 3, query 1, terms 4, landmarks 2, SMPL-H bwd 4, Adam 1, end 1)

    def heights(self):
        """get_smpl_height (recon_fit_base.py:818-828): extent of the vertices along y, [B]."""
        self.enqueue_forward()
        v = self.buf["verts"]
        return (v[:, :, 1].amax(1) - v[:, :, 1].amin(1)).clone()

    def run(self, weights: Dict[str, float], iter_for_betas, iter_for_pose, iter_for_kpts, steps_per_iter, max_iter, use_graph=True) -> bool:
        """The loop of recon_fit_behave.py:412-458.  Returns whether the early stop fired."""
        n_it = iter_for_betas + iter_for_kpts + iter_for_pose + max_iter
        sched = self.fitter.smpl_phase_schedule(iter_for_betas, iter_for_pose, iter_for_kpts, max_iter)
        rows = []
        for it, (phase, _) in enumerate(sched):
            decay = 1 if phase != "kpts" else it / 3
            row = [0.0] * 32
            for k, name in enumerate(SMPL_TERMS):
                row[k] = weights[name] / (1 + decay) if (name != "j2d" or phase == "kpts") else 0.0
            row[RC_LR0] = 0.02 if phase == "global" else 0.006
            row[RC_PHASE] = 0.0 if phase == "global" else 1.0
            row[RC_TOL] = 0.001
            row[RC_ESTOP] = 1.0 if it > 0.25 * max_iter + iter_for_betas + iter_for_pose else 0.0
            rows.append(row)
        with torch.cuda.device(self.device):
            self._upload_schedule(rows)
            self._start()
            stopped = False
            for it in range(n_it):
                if sched[it][1]:
                    self._new_optimizer()
                self._set_row(it)
                window = rows[it][RC_ESTOP] != 0.0
                for _ in range(steps_per_iter):
                    self._run(0, self._enqueue_step, use_graph)
                    if window and self._poll_stop():
                        stopped = True
                        break
                if stopped:
                    break
            torch.cuda.current_stream().synchronize()
            stopped = bool(self.ctrl[RC_STOP].item() != 0.0)
        return stopped

    def write_back(self, smpl):
        """The split parameters alias the caller's container in the reference (SMPLPyTorchWrapperBatchSplitParams.from_smpl wraps views of
        smpl.pose.data / betas.data / trans.data, lib_smpl/wrapper_pytorch.py:206-226), so every optimised value -- the other betas
        included -- ends up in it; copy_smpl_params (recon_fit_base.py:808-816) then re-copies a subset of the same storage."""
        b = self.buf
        with torch.no_grad():
            smpl.global_pose.copy_(b["pose"][:, :3]); smpl.body_pose.copy_(b["pose"][:, 3:66]); smpl.hand_pose.copy_(b["pose"][:, 66:])
            smpl.top_betas.copy_(b["betas"][:, :2]); smpl.other_betas.copy_(b["betas"][:, 2:]); smpl.trans.copy_(b["trans"])
            smpl.pose = torch.cat([smpl.global_pose, smpl.body_pose, smpl.hand_pose], 1)
            smpl.betas = torch.cat([smpl.top_betas, smpl.other_betas], 1)
        return smpl


class ObjectFitStep(_GraphLoop):
    TERMS = OBJ_TERMS
    PHASES = ("object only", "sil", "joint")

    def __init__(self, fitter, smpl_verts: torch.Tensor, data_dict, max_hist: int, inject_noise: bool, seed: int = 0):
        net = fitter.model
        dev = net.device
        self.fitter, self.net = fitter, net
        self._init_ctrl(dev, max_hist)
        to = dict(device=dev, dtype=torch.float32)
        self.obj_R, self.obj_t = data_dict["obj_R"], data_dict["obj_t"]              # the caller's leaf tensors, updated in place
        B = self.obj_R.shape[0]
        self.B = B
        if not (self.obj_R.is_cuda and self.obj_R.is_contiguous() and self.obj_t.is_contiguous() and self.obj_R.dtype == torch.float32):
            raise ValueError("obj_R / obj_t must be contiguous fp32 CUDA tensors")
        self.obj_s = data_dict["obj_s"].detach().to(**to).contiguous()
        objects = data_dict["objects"].detach().to(**to).contiguous()
        self.per_frame = 1 if objects.dim() == 3 else 0
        self.objects, self.N = objects, objects.shape[-2]
        self.occ = data_dict["occ_ratios"].detach().to(**to).contiguous()
        qd = data_dict["query_dict"]
        self.cc, self.bc = qd["crop_center"].to(**to).contiguous(), qd["body_center"].to(**to).contiguous()
        self.smpl_verts = smpl_verts.detach().to(**to).contiguous()
        self.sil = data_dict.get("silhouette")
        f = lambda *s: torch.empty(*s, **to)
        N = self.N
        b = self.buf = {}
        b["noise"], b["noise_used"], b["M"], b["R"] = torch.zeros(B, 9, **to), f(B, 9), f(B, 9), f(B, 9)
        b["object"], b["g_obj"], b["vals_df"], b["g_df"] = f(B, N, 3), f(B, N, 3), f(B, N), f(B, N, 3)
        b["gR"], b["gt"], b["gM"] = f(B, 9), f(B, 3), torch.zeros(B, 9, **to)
        b["t_init"] = self.obj_t.detach().clone()
        b["contact"] = torch.zeros(1, **to)
        self.m, self.v = torch.zeros(B, 12, **to), torch.zeros(B, 12, **to)
        self.inject = inject_noise
        self.seed = int(seed) & 0x7FFFFFFF
        self.pairs = None
        self._maps = net._maps
        self._div = (ctypes.c_double * 8)(max(B - 2, 1) * 3.0 * N, max(B - 1, 1) * 3.0 * N, B, B, 3.0 * B, float(B * N), 1.0, 1.0)
        if self.sil is not None:
            r = self.sil.renderer
            Vs, F, isz = self.sil.vertices.shape[0], r.faces.shape[0], r.image_size
            b["vsil"], b["g_vsil"] = f(B, Vs, 3), f(B, Vs, 3)
            b["faces_ndc"], b["g_faces"] = f(B, 2 * F, 9), f(B, 2 * F, 9)
            b["fidx"] = torch.empty(B, isz, isz, dtype=torch.int32, device=dev)
            b["alpha"], b["g_alpha"] = f(B, isz, isz), f(B, isz, isz)
            b["cull"] = torch.empty(_lib.load().vt_raster_cull_floats(B, F), **to)
            b["skip"] = torch.empty(_lib.load().vt_workspace_bytes_raster_bwd(B, isz), dtype=torch.uint8, device=dev)

    def _mutable_state(self):
        return [self.obj_R, self.obj_t, self.m, self.v, self.ctrl, self.acc, self.hist]

    # ---- pieces -------------------------------------------------------------------------------------------------------------------
    def enqueue_pose(self):
        """decopose_axis + transform_obj_verts with the noise of the current draw -> buf['R'], buf['object']."""
        b, B = self.buf, self.B
        _lib.call("vt_recon_obj_noise", P(self.obj_R), P(b["noise"]) if self.inject else None, B, P(self.ctrl), P(b["M"]), P(b["noise_used"]), S())
        _lib.call("vt_so3_project_fwd", P(b["M"]), B, P(b["R"]), S())
        _lib.call("vt_recon_obj_transform", P(self.objects), self.per_frame, P(b["R"]), P(self.obj_t), P(self.obj_s), B, self.N, P(b["object"]), S())

    def set_contact_pairs(self, pairs):
        """(h_idx, o_idx, h_off, o_off) from ``ReconFitterTriVisFull.contact_pairs`` or None; the human side is frozen -> gathered once."""
        self.pairs = pairs
        self.graphs.pop(2, None)
        if pairs is None:
            return
        h_idx, o_idx, h_off, o_off = pairs
        b, dev = self.buf, self.device
        self.h_idx, self.o_idx = h_idx.to(dev, torch.int64).contiguous(), o_idx.to(dev, torch.int64).contiguous()
        self.h_off, self.o_off = h_off.to(dev, torch.int32).contiguous(), o_off.to(dev, torch.int32).contiguous()
        nh, no = self.h_idx.numel(), self.o_idx.numel()
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        b["hs"], b["os"], b["g_hs"], b["g_os"] = f(nh, 3), f(no, 3), f(nh, 3), f(no, 3)
        b["nn_x"], b["nn_y"] = torch.empty(nh, dtype=torch.int32, device=dev), torch.empty(no, dtype=torch.int32, device=dev)
        _lib.call("vt_recon_gather_rows", P(self.smpl_verts), P(self.h_idx), nh, P(b["hs"]), S())

    def _enqueue_step(self, phase: int):
        b, B, N = self.buf, self.B, self.N
        ctrl, acc = P(self.ctrl), P(self.acc)
        self.enqueue_pose()
        if phase != 1:
            self.net.enqueue_query_losses(b["object"], self.cc, self.bc, 1, 0.8, None, b["vals_df"], b["g_df"], None, None, maps=self._maps)
            _lib.call("vt_recon_point_terms", P(b["object"]), B, N, 0, 1, 0, 1, 1, P(b["vals_df"]), P(b["g_df"]), P(self.occ), 5, 5, float(B * N),
                      None, None, -1, 0, 1.0, ctrl, P(b["g_obj"]), acc, S())
        else:
            # the reference still queries the network in the 'sil' phase (recon_fit_trivis_full.py:199-204) but none of its loss terms reads a
            # prediction: the launch is skipped
            _lib.call("vt_recon_point_terms", P(b["object"]), B, N, 0, 1, 0, 1, 1, None, None, None, -1, 0, 1.0, None, None, -1, 0, 1.0, ctrl,
                      P(b["g_obj"]), acc, S())
        contact = phase == 2 and self.pairs is not None
        if contact:
            nh, no, n_pairs = self.h_idx.numel(), self.o_idx.numel(), self.h_off.numel() - 1
            _lib.call("vt_recon_gather_rows", P(b["object"]), P(self.o_idx), no, P(b["os"]), S())
            _lib.call("vt_chamfer_fwd", P(b["hs"]), P(self.h_off), P(b["os"]), P(self.o_off), n_pairs, P(b["nn_x"]), P(b["nn_y"]), P(b["contact"]), S())
            _zero(b["g_hs"]); _zero(b["g_os"])
            _lib.call("vt_chamfer_bwd", P(b["hs"]), P(self.h_off), P(b["os"]), P(self.o_off), n_pairs, P(b["nn_x"]), P(b["nn_y"]),
                      ctypes.c_void_p(self.ctrl.data_ptr() + 4 * 6), P(b["g_hs"]), P(b["g_os"]), S())
            _lib.call("vt_recon_scatter_add_rows", P(b["g_os"]), P(self.o_idx), no, P(b["g_obj"]), S())
        _lib.call("vt_recon_obj_transform_bwd", P(self.objects), self.per_frame, P(b["g_obj"]), P(self.obj_s), B, N, 0, P(b["gR"]), P(b["gt"]), S())
        if phase == 1:
            sil, r = self.sil, self.sil.renderer
            Vs, F, isz = sil.vertices.shape[0], r.faces.shape[0], r.image_size
            _lib.call("vt_recon_obj_transform", P(sil.vertices), 0, P(b["R"]), P(self.obj_t), P(self.obj_s), B, Vs, P(b["vsil"]), S())
            _lib.call("vt_raster_fwd", P(b["vsil"]), P(r.faces), B, Vs, F, r.mode, P(r.K4), isz, P(b["faces_ndc"]), P(b["fidx"]), P(b["alpha"]), None,
                      P(b["cull"]), S())
            _lib.call("vt_recon_sil_loss", P(b["alpha"]), P(sil.keep_mask), P(sil.image_ref), P(self.occ), B, isz, ctrl, P(b["g_alpha"]), acc, S())
            _lib.call("vt_raster_bwd_ws", P(b["vsil"]), P(r.faces), B, Vs, F, r.mode, P(r.K4), isz, P(b["faces_ndc"]), P(b["fidx"]), P(b["alpha"]),
                      P(b["g_alpha"]), P(b["g_faces"]), P(b["g_vsil"]), P(b["skip"]), S())
            _lib.call("vt_recon_obj_transform_bwd", P(sil.vertices), 0, P(b["g_vsil"]), P(self.obj_s), B, Vs, 1, P(b["gR"]), P(b["gt"]), S())
        _lib.call("vt_recon_obj_small_terms", P(self.obj_t), P(b["t_init"]), P(self.obj_s), float(self.fitter.obj_scale), B, 1 if phase == 1 else 0,
                  ctrl, P(b["gt"]), acc, S())
        if phase != 2:                         # Adam([obj_t]) in the joint phase: the rotation gets no update
            _lib.call("vt_so3_project_bwd", P(b["M"]), P(b["gR"]), B, P(b["gM"]), S())
        _lib.call("vt_recon_adam_obj", P(self.obj_R), P(self.obj_t), P(b["gM"]), P(b["gt"]), P(self.m), P(self.v), B, ctrl, S())
        _lib.call("vt_recon_end_step", acc, len(OBJ_TERMS), self._div, 6 if contact else -1, P(b["contact"]) if contact else None, ctrl,
                  P(self.hist), self.max_hist, S())

    def step(self, phase: int, use_graph=True):
        self._run(phase, lambda: self._enqueue_step(phase), use_graph)
