"""One-time re-packing of a reference checkpoint (``CHORETriplaneVisibility.state_dict()``, 706 tensors, NCHW conv
weights; layout table in SURVEY.md section 5) into the layouts the sm_100a kernels read."""
from __future__ import annotations

from typing import Dict, Tuple

import torch

LO_SCALE = 2048.0          # csrc/common.cuh kLoScale
MMA_KC = 64                # csrc/conv_mma.cu MM_KC


def split_f16(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 -> (hi, lo) fp16 planes with x ~= hi + lo * 2^-11 (same rounding as the device-side split)."""
    x = x.float()
    if float(x.abs().max()) > 65504.0:
        raise ValueError("weight magnitude exceeds the fp16 range; the fp16x2 tensor-core path cannot represent it")
    hi = x.half()
    lo = ((x - hi.float()) * LO_SCALE).half()
    return hi.contiguous(), lo.contiguous()


def pack_conv(w: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Conv2d weight [Cout, Cin, k, k] -> {'ffma': fp32 [k*k, Cin, Cout], 'hi'/'lo': fp16 [k*k, Cout, Cin_pad]}."""
    cout, cin, kh, kw = w.shape
    assert kh == kw
    taps = w.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin).float()          # [tap, Cout, Cin]
    cin_pad = (cin + MMA_KC - 1) // MMA_KC * MMA_KC
    padded = torch.zeros(kh * kw, cout, cin_pad, dtype=torch.float32, device=w.device)
    padded[:, :, :cin] = taps
    hi, lo = split_f16(padded)
    return {"ffma": taps.transpose(1, 2).contiguous(), "hi": hi, "lo": lo, "ks": kh, "cin": cin, "cout": cout,
            "cin_pad": cin_pad}


def pack_stem(w: torch.Tensor) -> torch.Tensor:
    """Conv2d(cin, cout, 7, stride 2) weight [Cout, Cin, 7, 7] -> fp32 [49*Cin, Cout] (tap-major, then input channel)."""
    cout, cin = w.shape[:2]
    return w.permute(2, 3, 1, 0).reshape(49 * cin, cout).float().contiguous()


# ---- decoder packing (csrc/query.cu) ---------------------------------------------------------------------------
Q_K, Q_H = 616, 128
HEADS = ("df", "pca_predictor", "part_predictor", "center_predictor", "visib_predictor")   # output order df|pca|parts|centers|vis
HEAD_NOUT = (2, 9, 14, 3, 1)


def feature_permutation(c_im=256, c_tmpx=64, c_tt=32, c_tf=64) -> torch.Tensor:
    """internal index -> reference feature index.  Reference order (model/chore_triplane.py:139-151):
    im_feat | x,y,z-2.2 | tmpx | tri_tmpx r,b,t | tri_feat r | b | t ; internal order moves the 3 scalars to the end."""
    n_rest = c_tmpx + 3 * c_tt + 3 * c_tf
    ref = list(range(c_im)) + [c_im + 3 + i for i in range(n_rest)] + [c_im, c_im + 1, c_im + 2]
    return torch.tensor(ref, dtype=torch.long)


def pack_decoders(sd: Dict[str, torch.Tensor], device) -> torch.Tensor:
    """Five Conv1d(k=1) MLPs (model/chore.py:113-126) -> one fp32 buffer, k-major, first layer rows permuted/padded."""
    perm = feature_permutation()
    chunks = []
    for name, nout in zip(HEADS, HEAD_NOUT):
        w1 = sd[f"{name}.0.weight"][:, :, 0].float()                 # [128, 611]
        assert w1.shape == (Q_H, perm.numel()), f"{name}.0.weight has shape {tuple(w1.shape)}"
        w1p = torch.zeros(Q_K, Q_H)
        w1p[: perm.numel()] = w1[:, perm].t().cpu()
        chunks += [w1p.reshape(-1), sd[f"{name}.0.bias"].float().cpu()]
        for idx in (2, 4):
            chunks += [sd[f"{name}.{idx}.weight"][:, :, 0].float().t().contiguous().cpu().reshape(-1),
                       sd[f"{name}.{idx}.bias"].float().cpu()]
        w4 = torch.zeros(Q_H, 16)
        w4[:, :nout] = sd[f"{name}.6.weight"][:, :, 0].float().t().cpu()
        b4 = torch.zeros(16)
        b4[:nout] = sd[f"{name}.6.bias"].float().cpu()
        chunks += [w4.reshape(-1), b4]
    return torch.cat(chunks).contiguous().to(device)


Q_KB = 640


def pack_decoders_bwd(sd: Dict[str, torch.Tensor], device) -> torch.Tensor:
    """Transposed copies for the analytic backward pass (csrc/query.cu): per head W1b [128][640] (feature columns in the
    internal order, zero padded), W2b, W3b [128][128] and W4b [16][128] -- i.e. the torch [out][in] layout."""
    perm = feature_permutation()
    chunks = []
    for name, nout in zip(HEADS, HEAD_NOUT):
        w1b = torch.zeros(Q_H, Q_KB)
        w1b[:, : perm.numel()] = sd[f"{name}.0.weight"][:, :, 0].float().cpu()[:, perm]
        chunks.append(w1b.reshape(-1))
        for idx in (2, 4):
            chunks.append(sd[f"{name}.{idx}.weight"][:, :, 0].float().cpu().contiguous().reshape(-1))
        w4b = torch.zeros(16, Q_H)
        w4b[:nout] = sd[f"{name}.6.weight"][:, :, 0].float().cpu()
        chunks.append(w4b.reshape(-1))
    return torch.cat(chunks).contiguous().to(device)


def feature_permutation_tc() -> torch.Tensor:
    """tensor-core kernel feature index (csrc/query_tc.cu, 640 padded) -> reference feature index, -1 for zero padding:
    im_feat 256 | tmpx 64 | tri_feat r, b, t 3x64 | tri_tmpx r 32, b 32 | tri_tmpx t 32, x, y, z-2.2, 29 zeros."""
    ref = list(range(256)) + list(range(259, 323)) + list(range(419, 611)) + list(range(323, 387)) + list(range(387, 419)) + [256, 257, 258]
    return torch.tensor(ref + [-1] * (640 - len(ref)), dtype=torch.long)


def split_f16_unscaled(x: torch.Tensor):
    """x ~= hi + lo with lo NOT rescaled (single-accumulator scheme of csrc/query_tc.cu)."""
    x = x.float()
    if float(x.abs().max()) > 65504.0:
        raise ValueError("weight magnitude exceeds the fp16 range")
    hi = x.half()
    return hi.contiguous(), (x - hi.float()).half().contiguous()


def pack_decoders_tc(sd: Dict[str, torch.Tensor], device):
    """fp16 hi/lo planes for the tcgen05 decoder kernel: W1 [5*128, 640] (rows = head, unit; columns in the kernel's feature
    order) and W2|W3 [2*5*128, 128] (layer-major, then head) -- torch's [out][in] layout, i.e. K-major B operands."""
    perm = feature_permutation_tc()
    w1 = torch.zeros(5 * Q_H, 640)
    w23 = torch.zeros(2 * 5 * Q_H, Q_H)
    valid = perm >= 0
    for h, name in enumerate(HEADS):
        w = sd[f"{name}.0.weight"][:, :, 0].float().cpu()
        w1[h * Q_H:(h + 1) * Q_H][:, valid] = w[:, perm[valid]]
        for li, idx in enumerate((2, 4)):
            w23[(li * 5 + h) * Q_H:(li * 5 + h + 1) * Q_H] = sd[f"{name}.{idx}.weight"][:, :, 0].float().cpu()
    w1h, w1l = split_f16_unscaled(w1)
    w2h, w2l = split_f16_unscaled(w23)
    return tuple(t.to(device) for t in (w1h, w1l, w2h, w2l))


def pack_decoders_tc_bwd(sd: Dict[str, torch.Tensor], device):
    """Transposed fp16 hi/lo planes for the tcgen05 backward kernel (csrc/query_bwd_tc.cu): W2^T | W3^T [2*5*128, 128] (rows = input
    unit j, columns = output unit k: the B operand of g_in[j] = sum_k g_out[k] W[k][j]) and W1^T [5*640, 128] (rows = feature in the
    kernel's order, zero rows for the padding)."""
    perm = feature_permutation_tc()
    valid = perm >= 0
    w23t = torch.zeros(2 * 5 * Q_H, Q_H)
    w1t = torch.zeros(5 * 640, Q_H)
    for h, name in enumerate(HEADS):
        w = sd[f"{name}.0.weight"][:, :, 0].float().cpu()                     # [128 out, 611 in]
        blk = torch.zeros(640, Q_H)
        blk[valid] = w[:, perm[valid]].t()
        w1t[h * 640:(h + 1) * 640] = blk
        for li, idx in enumerate((2, 4)):
            w23t[(li * 5 + h) * Q_H:(li * 5 + h + 1) * Q_H] = sd[f"{name}.{idx}.weight"][:, :, 0].float().cpu().t()
    a, b = split_f16_unscaled(w23t)
    c, d = split_f16_unscaled(w1t)
    return tuple(t.to(device) for t in (a, b, c, d))
