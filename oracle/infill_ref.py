"""CPU restatement of HVOP-Net -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows the reference's
  * ``PositionEmbeddingSine_1D.forward`` (model/transformers/posi_embed.py:35-66; float32 table, normalised positions, temperature 10000,
    exponent 2 i / N with i the feature index -- not i // 2),
  * ``TransformerEncoderLayer.forward_pre`` (model/transformers/former_deci.py:78-93) -- every layer is pre-norm because ``TransformerV2``
    constructs it with ``pre_norm=True`` (:139-143); the ``pre_norm`` option only adds the closing LayerNorm (:144) -- with
    ``nn.MultiheadAttention`` written out (packed in-projection, per-head scaled dot product, key padding mask -> -inf, out-projection),
  * ``ConditionalMInfiller.forward`` (model/infill/mfiller_cond.py:78-104) and ``make_predictor`` (:57-73),
  * the autoregressive clip loop ``MotionInfillAutoreg.test`` (interp/test_infill_autoreg.py:78-163) with
    ``CondMotionInfillAutoreg.model_forward`` (interp/test_cinfill_autoreg.py:32-51) for ``obj_dim == 6``,
  * ``rot6d_to_rotmat`` (utils/geometry_utils.py:63-77) and the stored layout ``obj_angles = R^T`` (interp/test_infiller.py:134).
float64 throughout except the positional table.  Pinned by tests/golden/infill_small.npz, produced by tests/golden/make_golden.py
(``--only infill``) from the reference's own modules and its own ``test()`` loop.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def _get(opt, k):
    return opt[k] if isinstance(opt, dict) else getattr(opt, k)


def position_embedding(L: int, D: int) -> torch.Tensor:
    """[L, D] float32 (identical for every clip of the batch)."""
    n = D // 2
    pos = torch.arange(0, L, dtype=torch.float32)
    pos = pos / (pos[-1:] + 1e-6) * (2 * math.pi)
    dim_t = 10000 ** (2 * torch.arange(n, dtype=torch.float32) / n)
    pe = torch.zeros(L, D)
    ang = pos[:, None] / dim_t
    if 2 * n != D:
        pe[:, :-1][:, 0::2] = torch.sin(ang)
    else:
        pe[:, 0::2] = torch.sin(ang)
    pe[:, 1::2] = torch.cos(ang)
    return pe


def _act(name: str):
    return {"gelu": F.gelu, "relu": F.relu, "leaky_relu": F.leaky_relu}[name]


def encoder_layer(sd, p: str, x: torch.Tensor, key_mask, pos: torch.Tensor, heads: int, act) -> torch.Tensor:
    """x [B, T, D] float64; key_mask [B, T] bool (True = ignored) or None."""
    B, T, D = x.shape
    dh = D // heads
    w = lambda k: sd[p + k].double()
    h = F.layer_norm(x, (D,), w("norm1.weight"), w("norm1.bias"), 1e-5)
    qk_in = h + pos.double()
    Wi, bi = w("self_attn.in_proj_weight"), w("self_attn.in_proj_bias")
    q = (qk_in @ Wi[:D].T + bi[:D]) / math.sqrt(dh)
    k = qk_in @ Wi[D:2 * D].T + bi[D:2 * D]
    v = h @ Wi[2 * D:].T + bi[2 * D:]
    split = lambda t: t.reshape(B, T, heads, dh).permute(0, 2, 1, 3)
    s = split(q) @ split(k).transpose(-1, -2)                                      # [B, heads, T, T]
    if key_mask is not None:
        s = s.masked_fill(key_mask[:, None, None, :], float("-inf"))
    a = torch.softmax(s, -1) @ split(v)
    a = a.permute(0, 2, 1, 3).reshape(B, T, D)
    x = x + a @ w("self_attn.out_proj.weight").T + w("self_attn.out_proj.bias")
    h2 = F.layer_norm(x, (D,), w("norm2.weight"), w("norm2.bias"), 1e-5)
    f = act(h2 @ w("linear1.weight").T + w("linear1.bias"))
    return x + f @ w("linear2.weight").T + w("linear2.bias")


def transformer(sd, opt, name: str, x: torch.Tensor, key_mask) -> torch.Tensor:
    B, T, D = x.shape
    pos = position_embedding(T, D)
    for i in range(_get(opt, "num_layers_" + name)):
        x = encoder_layer(sd, f"encoder_{name}.encoder.layers.{i}.", x, key_mask, pos, _get(opt, "num_heads_" + name), _act(_get(opt, "activation_" + name)))
    if _get(opt, "pre_norm_" + name):
        x = F.layer_norm(x, (D,), sd[f"encoder_{name}.encoder.norm.weight"].double(), sd[f"encoder_{name}.encoder.norm.bias"].double(), 1e-5)
    return x


def cond_infiller_forward(sd, opt, data_smpl, mask_smpl, data_obj, mask_obj) -> torch.Tensor:
    """[B, T, dim_smpl], [B, T] bool, [B, T, dim_obj], [B, T] bool -> [B, T, out_dim] float64."""
    ds, do = torch.as_tensor(data_smpl).double(), torch.as_tensor(data_obj).double()
    fs = transformer(sd, opt, "smpl", ds @ sd["feat_proj_smpl.weight"].double().T + sd["feat_proj_smpl.bias"].double(), torch.as_tensor(mask_smpl).bool())
    fo = transformer(sd, opt, "obj", do @ sd["feat_proj_obj.weight"].double().T + sd["feat_proj_obj.bias"].double(), torch.as_tensor(mask_obj).bool())
    x = transformer(sd, opt, "joint", torch.cat([fs, fo], -1), None)
    n = len(_get(opt, "hidden_dims")) + 1
    for i in range(n):
        x = x @ sd[f"predictor.{2 * i}.weight"].double().T + sd[f"predictor.{2 * i}.bias"].double()
        if i < n - 1:
            x = F.leaky_relu(x)
    return x


def rot6d_to_rotmat(x: torch.Tensor) -> torch.Tensor:
    x = x.reshape(-1, 3, 2)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = F.normalize(a1)
    b2 = F.normalize(a2 - (b1 * a2).sum(-1, keepdim=True) * b1)
    return torch.stack((b1, b2, torch.linalg.cross(b1, b2)), -1)


def autoreg_infill(sd, opt, rot6d_smpl, trans_smpl, rot6d_obj, trans_obj, occ_ratios, occ_thres=0.5, clip_len=None, window=30, init_thres=0.5):
    """The clip loop for obj_dim 6.  Returns (obj_angles [L,3,3] = R^T, obj_trans [L,3], rot6d_out [L,6]) or None when the first clip has
    fewer than `window` visible frames (the reference then saves its input unchanged)."""
    clip_len = clip_len or _get(opt, "clip_len")
    L = rot6d_obj.shape[0]
    rot6d_out = np.zeros((L, 6), np.float64)

    def forward(data, mask):
        data = data.copy()
        data[:, -6:] = data[:, -6:] * (1 - mask.astype(float)[:, None])
        d = torch.from_numpy(data[None]).float()                                   # the reference feeds float32 to the network
        m = torch.from_numpy(mask[None])
        return cond_infiller_forward(sd, opt, d[:, :, :-6], torch.zeros_like(m), d[:, :, -6:], m)[0].numpy()

    s, e = 0, clip_len
    mask = occ_ratios[s:e] < init_thres
    if np.sum(~mask) < window:
        return None
    data = np.concatenate([rot6d_smpl[s:e], trans_smpl[s:e], rot6d_obj[s:e]], 1).astype(np.float64)
    rot6d_out[s:e] = forward(data, mask).astype(np.float32)                        # pred is float32 in the reference
    for idx in range(0, L - clip_len + 1 + window, window):
        s, e = idx, idx + clip_len
        data = np.concatenate([rot6d_smpl[s:e], trans_smpl[s:e], rot6d_obj[s:e]], 1).astype(np.float64)
        data[:window, -6:] = rot6d_out[s:s + window]
        mask = occ_ratios[s:e] < occ_thres
        mask[:window] = False
        rot6d_out[s + window:e] = forward(data, mask)[window:].astype(np.float32)
    R = rot6d_to_rotmat(torch.from_numpy(rot6d_out))
    return R.transpose(1, 2).numpy(), np.array(trans_obj, copy=True), rot6d_out
