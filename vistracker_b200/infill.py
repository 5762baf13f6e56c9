"""HVOP-Net on B200: drop-ins for ``ConditionalMInfiller`` (model/infill/mfiller_cond.py:17-104) and for the
numeric core of ``CondMotionInfillAutoreg`` (interp/test_cinfill_autoreg.py:23-51 over interp/test_infill_autoreg.py:34-174).

The reference runs one 180-frame clip at a time through ``nn.MultiheadAttention`` layers (about 200 small launches per clip), with a numpy
round trip between clips because clip i+1 is seeded with the prediction of clip i.  Here the trajectory, the running prediction and every
intermediate stay in device memory; a clip is 19 launches of libvistracker_sm100a.so (vt_infill_head / vt_infill_attn / vt_infill_tail /
vt_infill_mlp) plus the clip gather and commit kernels; the whole sequence is enqueued without a host synchronisation and, by default,
replayed as one CUDA graph per sequence length.  File IO (joblib packs) stays with the caller.
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib

P, S = _lib.ptr, _lib.stream_ptr


def _on_own_device(fn):
    """Run a method with the module's device current: ``_lib.stream_ptr()``, the kernel launches and CUDA-graph capture all act on the
    CURRENT device (a multi-GPU process may not have called ``torch.cuda.set_device``)."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **kw):
        with torch.cuda.device(self.device):
            return fn(self, *a, **kw)
    return wrapped
_ACT = {"gelu": 0, "relu": 1, "leaky_relu": 2}


def default_infill_options() -> dict:
    """The values of config/cmf-k4-lrot.json that shape the network and the clip loop (the model scripts/demo.sh:33 runs)."""
    return dict(clip_len=180, dim_smpl=147, dim_obj=6, out_dim=6, obj_repre="6d",
                num_layers_smpl=2, d_model_smpl=128, num_heads_smpl=4, dim_forward_smpl=256, pre_norm_smpl=False, activation_smpl="gelu",
                num_layers_obj=2, d_model_obj=32, num_heads_obj=2, dim_forward_obj=64, pre_norm_obj=False, activation_obj="gelu",
                num_layers_joint=4, num_heads_joint=1, dim_forward_joint=256, pre_norm_joint=False, activation_joint="gelu",
                hidden_dims=[32])


def position_embedding(L: int, D: int) -> torch.Tensor:
    """``PositionEmbeddingSine_1D(D // 2, normalize=True, total_feat_dim=D)(B, L)[:, 0]`` (model/transformers/posi_embed.py:35-66), built in
    float32 with the reference's operation order so that the table is the same to the last bit."""
    n = D // 2
    t = torch.arange(0, L, dtype=torch.float32)
    t = t / (t[-1:] + 1e-6) * (2 * math.pi)
    freq = torch.arange(n, dtype=torch.float32)
    freq = 10000 ** (2 * torch.div(freq, 1) / n)
    table = torch.zeros(L, D)
    table[:, 0:2 * n:2] = torch.sin(t[:, None] / freq)
    table[:, 1::2] = torch.cos(t[:, None] / freq)
    return table


class _Encoder:
    """One ``TransformerV2`` (model/transformers/former_deci.py:132-175): packed layers and the launch sequence."""

    def __init__(self, sd, name: str, opt, D: int, device):
        g = (lambda k: opt[k]) if isinstance(opt, dict) else (lambda k: getattr(opt, k))
        self.D, self.F, self.heads = D, int(g("dim_forward_" + name)), int(g("num_heads_" + name))
        act = g("activation_" + name)
        if act not in _ACT:
            raise RuntimeError(f"activation should be relu/gelu/leaky_relu, not {act}.")          # glu changes the layer widths
        self.act = _ACT[act]
        if D % self.heads:
            raise RuntimeError(f"embed_dim {D} must be divisible by num_heads {self.heads}")
        n_floats = _lib.load().vt_infill_layer_pack_floats(D, self.F)
        self.layers = []
        for i in range(int(g("num_layers_" + name))):
            p = f"encoder_{name}.encoder.layers.{i}."
            w = lambda k: sd[p + k].detach().float().cpu()
            parts = [w("norm1.weight"), w("norm1.bias"), w("self_attn.in_proj_weight").t().contiguous().reshape(-1), w("self_attn.in_proj_bias"),
                     w("self_attn.out_proj.weight").t().contiguous().reshape(-1), w("self_attn.out_proj.bias"), w("norm2.weight"), w("norm2.bias"),
                     w("linear1.weight").t().contiguous().reshape(-1), w("linear1.bias"), w("linear2.weight").t().contiguous().reshape(-1), w("linear2.bias")]
            if tuple(sd[p + "self_attn.in_proj_weight"].shape) != (3 * D, D) or tuple(sd[p + "linear1.weight"].shape) != (self.F, D):
                raise RuntimeError(f"{p}: checkpoint shapes do not match d_model {D} / dim_forward {self.F}")
            buf = torch.cat(parts).contiguous()
            assert buf.numel() == n_floats
            self.layers.append(buf.to(device))
        if not self.layers:
            raise RuntimeError(f"encoder_{name} has no layers")
        self.final_ln = None
        if g("pre_norm_" + name):
            self.final_ln = torch.cat([sd[f"encoder_{name}.encoder.norm.weight"].float().cpu(), sd[f"encoder_{name}.encoder.norm.bias"].float().cpu()]).to(device)

    def run(self, data, proj, x, n_tok, T, pos, key_mask, qkv, attn, y, y_ld):
        """data/proj: the projected input (or None: x already holds the stream); result rows -> y[n_tok][y_ld]."""
        D, F, H = self.D, self.F, self.heads
        if data is not None:
            _lib.call("vt_infill_head", P(data), data.shape[-1], data.shape[-1], P(proj), P(x), D, n_tok, T, D, F, H, P(self.layers[0]), P(pos), P(qkv), S())
        else:
            _lib.call("vt_infill_head", None, 0, 0, None, P(x), D, n_tok, T, D, F, H, P(self.layers[0]), P(pos), P(qkv), S())
        for i, layer in enumerate(self.layers):
            _lib.call("vt_infill_attn", P(qkv), P(key_mask) if key_mask is not None else None, n_tok // T, T, D, H, P(attn), S())
            last = i == len(self.layers) - 1
            nxt = None if last else self.layers[i + 1]
            _lib.call("vt_infill_tail", P(x), D, P(attn), n_tok, T, D, F, H, self.act, P(layer),
                      P(self.final_ln) if (last and self.final_ln is not None) else None,
                      P(y) if last else P(x), y_ld if last else D, P(nxt) if nxt is not None else None, F, P(pos), P(qkv), S())


class ConditionalMInfiller:
    """``ConditionalMInfiller`` (model/infill/mfiller_cond.py): SMPL encoder, masked object encoder, joint encoder, predictor; inference only."""

    def __init__(self, opt=None, device=None):
        self.opt = opt if opt is not None else default_infill_options()
        self.device = torch.device(device) if device is not None else torch.device("cuda", 0)
        if self.device.type != "cuda":
            raise RuntimeError("vistracker_b200 has no CPU path: ConditionalMInfiller needs a CUDA device")
        _lib.load()
        self._loaded = False
        self._pos: Dict[tuple, torch.Tensor] = {}
        self._ws: Dict[tuple, dict] = {}
        self._pinned = set()            # workspaces a captured CUDA graph points into: never dropped

    def _g(self, k):
        return self.opt[k] if isinstance(self.opt, dict) else getattr(self.opt, k)

    def eval(self):
        return self

    def to(self, device):
        if torch.device(device) != self.device:
            raise RuntimeError("move the checkpoint, not the module: construct ConditionalMInfiller(device=...)")
        return self

    def load_state_dict(self, sd, strict: bool = True):
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}      # DataParallel checkpoints (trainer/train_utils.py)
        from .synth import infill_spec
        want = [k for k, _, _ in infill_spec(self.opt)]
        missing = [k for k in want if k not in sd]
        if missing or (strict and len(sd) != len(want)):
            extra = [k for k in sd if k not in want]
            raise RuntimeError(f"Error(s) in loading state_dict for ConditionalMInfiller: missing {missing[:4]}, unexpected {extra[:4]}")
        dev = self.device
        self.Ds, self.Do = int(self._g("d_model_smpl")), int(self._g("d_model_obj"))
        self.Dj = self.Ds + self.Do
        self.enc_smpl = _Encoder(sd, "smpl", self.opt, self.Ds, dev)
        self.enc_obj = _Encoder(sd, "obj", self.opt, self.Do, dev)
        self.enc_joint = _Encoder(sd, "joint", self.opt, self.Dj, dev)
        pk = lambda n: torch.cat([sd[n + ".weight"].float().cpu().t().contiguous().reshape(-1), sd[n + ".bias"].float().cpu()]).to(dev)
        self.proj_smpl, self.proj_obj = pk("feat_proj_smpl"), pk("feat_proj_obj")
        self.mlp_dims = [self.Dj] + [int(h) for h in self._g("hidden_dims")] + [int(self._g("out_dim"))]
        if len(self.mlp_dims) - 1 > 5:
            raise RuntimeError("the predictor kernel takes at most 5 Linear layers")
        self.mlp_pack = torch.cat([pk(f"predictor.{2 * i}") for i in range(len(self.mlp_dims) - 1)])
        self._dims_c = (ctypes.c_int * len(self.mlp_dims))(*self.mlp_dims)
        self._loaded = True
        return self

    def _pos_table(self, T: int, D: int) -> torch.Tensor:
        if (T, D) not in self._pos:
            self._pos[(T, D)] = position_embedding(T, D).to(self.device)
        return self._pos[(T, D)]

    def _workspace(self, B: int, T: int, pin: bool = False) -> dict:
        key = (B, T)
        if pin:
            self._pinned.add(key)
        if key not in self._ws:
            if len(self._ws) > 8:
                for k in [k for k in self._ws if k not in self._pinned]:
                    del self._ws[k]
            n, dev = B * T, self.device
            e = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
            self._ws[key] = dict(xs=e(n, self.Ds), xo=e(n, self.Do), xj=e(n, self.Dj), qkv=e(n, 3 * self.Dj), qkv2=e(n, 3 * self.Do), attn=e(n, self.Dj),
                                 attn2=e(n, self.Do), side=torch.cuda.Stream(device=dev))
        return self._ws[key]

    @_on_own_device
    def forward_into(self, data_smpl, mask_smpl, data_obj, mask_obj, pred, pin: bool = False):
        """All arguments device tensors: data [B,T,*] float32 contiguous, masks [B,T] uint8/bool or None, pred [B,T,out_dim].  ``pin`` keeps
        the (B, T) workspace alive for the lifetime of the module (callers that capture the launches into a CUDA graph)."""
        if not self._loaded:
            raise RuntimeError("load_state_dict first")
        B, T = data_smpl.shape[:2]
        n = B * T
        ws = self._workspace(B, T, pin)
        main = torch.cuda.current_stream(self.device)
        side = ws["side"]
        side.wait_stream(main)
        with torch.cuda.stream(side):                                                  # the object branch runs beside the SMPL branch
            self.enc_obj.run(data_obj, self.proj_obj, ws["xo"], n, T, self._pos_table(T, self.Do), mask_obj, ws["qkv2"], ws["attn2"],
                             ws["xj"][:, self.Ds:], self.Dj)
        self.enc_smpl.run(data_smpl, self.proj_smpl, ws["xs"], n, T, self._pos_table(T, self.Ds), mask_smpl, ws["qkv"], ws["attn"], ws["xj"], self.Dj)
        main.wait_stream(side)
        self.enc_joint.run(None, None, ws["xj"], n, T, self._pos_table(T, self.Dj), None, ws["qkv"], ws["attn"], ws["xj"], self.Dj)
        _lib.call("vt_infill_mlp", P(ws["xj"]), self.Dj, n, len(self.mlp_dims) - 1, self._dims_c, P(self.mlp_pack), P(pred), self.mlp_dims[-1], S())
        return pred

    def forward(self, data_smpl, mask_smpl, data_obj, mask_obj):
        if not self._loaded:
            raise RuntimeError("load_state_dict first")
        dev = self.device
        ds = torch.as_tensor(data_smpl).to(dev, torch.float32).contiguous()
        do = torch.as_tensor(data_obj).to(dev, torch.float32).contiguous()
        if ds.dim() != 3 or do.shape[:2] != ds.shape[:2] or ds.shape[2] != self._g("dim_smpl") or do.shape[2] != self._g("dim_obj"):
            raise RuntimeError(f"expected (B, T, {self._g('dim_smpl')}) and (B, T, {self._g('dim_obj')}), got {tuple(ds.shape)} and {tuple(do.shape)}")
        mk = lambda m: None if m is None else torch.as_tensor(m).to(dev).to(torch.uint8).contiguous()
        pred = torch.empty(ds.shape[0], ds.shape[1], self.mlp_dims[-1], device=dev)
        return self.forward_into(ds, mk(mask_smpl), do, mk(mask_obj), pred)

    __call__ = forward


class CondMotionInfillAutoreg:
    """The autoregressive in-filling of one sequence (interp/test_infill_autoreg.py:78-163 with the conditional model_forward of
    interp/test_cinfill_autoreg.py:32-51, ``obj_repre == '6d'``): clips of ``clip_len`` frames every ``window`` frames, the first ``window``
    frames of a clip seeded with the previous prediction, frames whose visibility is below the threshold masked and predicted."""

    def __init__(self, model: ConditionalMInfiller, clip_len: Optional[int] = None, window: int = 30, init_thres: float = 0.5, use_graph: Optional[bool] = None):
        self.model, self.device = model, model.device
        self.clip_len = int(clip_len if clip_len is not None else model._g("clip_len"))
        self.window, self.init_thres = int(window), float(init_thres)
        if model._g("dim_obj") != 6 or model._g("out_dim") != 6:
            raise RuntimeError("only the rotation-only model (obj_repre '6d', out_dim 6) is built")
        self.use_graph = (os.environ.get("VT_INFILL_GRAPH", "1") != "0") if use_graph is None else use_graph
        self._plans: Dict[int, dict] = {}

    def clip_plan(self, L: int):
        """[(start, T, n_ctx)]: the first clip, then the strided ones (a last clip may be shorter than clip_len)."""
        plan = [(0, min(self.clip_len, L), 0)]
        for idx in range(0, L - self.clip_len + 1 + self.window, self.window):
            T = min(self.clip_len, L - idx)
            plan.append((idx, T, min(self.window, T)))
        return plan

    def _masks(self, occ: np.ndarray, occ_thres: float, plan):
        rows = np.zeros((len(plan), self.clip_len), np.uint8)
        for i, (s, T, n_ctx) in enumerate(plan):
            m = occ[s:s + T] < (self.init_thres if i == 0 else occ_thres)
            m[:n_ctx] = False
            rows[i, :T] = m
        return rows

    def _buffers(self, L: int) -> dict:
        if L not in self._plans:
            if len(self._plans) > 4:
                self._plans.clear()
            dev, C = self.device, self.clip_len
            e = lambda *s, dt=torch.float32: torch.zeros(*s, device=dev, dtype=dt)
            plan = self.clip_plan(L)
            self._plans[L] = dict(plan=plan, rs=e(L, 144), ts=e(L, 3), ro=e(L, 6), out=e(L, 6), masks=e(len(plan), C, dt=torch.uint8), zero=e(C, dt=torch.uint8),
                                  ds=e(C, 147), do=e(C, 6), pred=e(C, 6), angles=e(L, 3, 3), graph=None)
        return self._plans[L]

    def _enqueue(self, b: dict, L: int):
        m = self.model
        b["out"].zero_()
        for i, (s, T, n_ctx) in enumerate(b["plan"]):
            ds, do, pred = b["ds"][:T], b["do"][:T], b["pred"][:T]
            _lib.call("vt_infill_pack_clip", P(b["rs"]), P(b["ts"]), P(b["ro"]), P(b["out"]), P(b["masks"][i]), L, s, T, n_ctx, P(ds), P(do), S())
            m.forward_into(ds[None], b["zero"][None, :T], do[None], b["masks"][i][None, :T], pred[None], pin=self.use_graph)
            _lib.call("vt_infill_commit_clip", P(pred), L, s, n_ctx, T, P(b["out"]), S())
        _lib.call("vt_smooth_rot6d_to_rotmat", P(b["out"]), L, 1, P(b["angles"]), S())          # stored as R^T (interp/test_infiller.py:134)

    @_on_own_device
    def infill(self, rot6d_smpl, trans_smpl, rot6d_obj, trans_obj, occ_ratios, occ_thres: float = 0.5):
        """rot6d_smpl [L,144], trans_smpl [L,3], rot6d_obj [L,6], trans_obj [L,3] (tensors or arrays), occ_ratios [L] (visible fraction, the
        first column of ``neural_visibility``).  Returns what ``save_output`` stores: ``obj_angles`` [L,3,3] (= R^T), ``obj_trans`` (a copy of the
        input: the rotation-only model leaves it), ``obj_scales`` ones, plus ``rot6d`` -- or ``None`` when the first clip has fewer than
        ``window`` visible frames (the reference then writes its input back unchanged)."""
        dev = self.device
        occ = occ_ratios.detach().cpu().numpy() if torch.is_tensor(occ_ratios) else np.asarray(occ_ratios)
        L = int(occ.shape[0])
        t = lambda a, d: torch.as_tensor(a).to(dev, torch.float32).reshape(L, d)
        b = self._buffers(L)
        masks = self._masks(occ, occ_thres, b["plan"])
        T0 = b["plan"][0][1]
        if int((masks[0, :T0] == 0).sum()) < self.window:
            return None
        b["rs"].copy_(t(rot6d_smpl, 144)); b["ts"].copy_(t(trans_smpl, 3)); b["ro"].copy_(t(rot6d_obj, 6))
        b["masks"].copy_(torch.from_numpy(masks), non_blocking=False)
        if self.use_graph:
            if b["graph"] is None:
                self._enqueue(b, L)                                                     # warm-up: smem attributes, workspaces, side streams
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue(b, L)
                b["graph"] = g
            b["graph"].replay()
        else:
            self._enqueue(b, L)
        return {"obj_angles": b["angles"].clone(), "obj_trans": t(trans_obj, 3).clone(), "obj_scales": torch.ones(L, device=dev), "rot6d": b["out"].clone()}
