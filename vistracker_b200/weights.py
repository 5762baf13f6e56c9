"""One-time re-packing of a reference checkpoint (``CHORETriplaneVisibility.state_dict()``, 706 tensors, NCHW conv weights; layout table
in SURVEY.md section 5) into the layouts the sm_100a kernels read.  The packing itself is host C behind the C ABI
(``vt_pack_weights_*``, csrc/pack.cu) so that a consumer without Python can load a checkpoint too; this module hands it the tensors of a
state dict and uploads the results.  (tests/test_pack_weights.py holds an independent torch statement of every layout.)"""
from __future__ import annotations

import ctypes
from typing import Dict

import torch

from . import _lib

LO_SCALE = 2048.0          # csrc/common.cuh kLoScale
MMA_KC = 64                # csrc/conv_mma.cu MM_KC
Q_K, Q_H, Q_KB = 616, 128, 640
HEADS = ("df", "pca_predictor", "part_predictor", "center_predictor", "visib_predictor")   # output order df|pca|parts|centers|vis
HEAD_NOUT = (2, 9, 14, 3, 1)


def _host32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to("cpu", torch.float32).contiguous()


def _hp(t: torch.Tensor):
    return ctypes.c_void_p(t.data_ptr())


def pack_conv(w: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Conv2d weight [Cout, Cin, k, k] -> {'ffma': fp32 [k*k, Cin, Cout], 'hi'/'lo': fp16 [k*k, Cout, Cin_pad]} (``vt_pack_weights_conv``);
    results live where ``w`` lives."""
    cout, cin, kh, kw = w.shape
    assert kh == kw
    wh = _host32(w)
    cin_pad = int(_lib.load().vt_conv_cin_pad(cin))
    ffma = torch.empty(kh * kw, cin, cout, dtype=torch.float32)
    hi = torch.empty(kh * kw, cout, cin_pad, dtype=torch.float16)
    lo = torch.empty_like(hi)
    try:
        _lib.call("vt_pack_weights_conv", _hp(wh), cout, cin, kh, _hp(ffma), _hp(hi), _hp(lo))
    except RuntimeError as e:
        raise ValueError(str(e)) from None       # a weight beyond the fp16 range: the fp16 x 2 tensor-core path cannot represent it
    return {"ffma": ffma.to(w.device), "hi": hi.to(w.device), "lo": lo.to(w.device), "ks": kh, "cin": cin, "cout": cout, "cin_pad": cin_pad}


def pack_stem(w: torch.Tensor) -> torch.Tensor:
    """Conv2d(cin, cout, 7, stride 2) weight [Cout, Cin, 7, 7] -> fp32 [49*Cin, Cout] (tap-major, then input channel)."""
    cout, cin = w.shape[:2]
    out = torch.empty(49 * cin, cout, dtype=torch.float32)
    wh = _host32(w)
    _lib.call("vt_pack_weights_stem", _hp(wh), cout, cin, _hp(out))
    return out.to(w.device)


def _decoder_tables(sd: Dict[str, torch.Tensor]):
    """(keep-alive tensors, weight pointer table, bias pointer table) in the order vt_pack_weights_decoders reads: head-major, 4 layers."""
    ws, bs = [], []
    for name, nout in zip(HEADS, HEAD_NOUT):
        for li, idx in enumerate((0, 2, 4, 6)):
            w = _host32(sd[f"{name}.{idx}.weight"][:, :, 0])
            want = (Q_H if li < 3 else nout, 611 if li == 0 else Q_H)
            assert tuple(w.shape) == want, f"{name}.{idx}.weight has shape {tuple(w.shape)}, expected {want}"
            ws.append(w)
            bs.append(_host32(sd[f"{name}.{idx}.bias"]))
    wt = (ctypes.c_void_p * 20)(*(t.data_ptr() for t in ws))
    bt = (ctypes.c_void_p * 20)(*(t.data_ptr() for t in bs))
    return (ws, bs), wt, bt


def pack_decoders(sd: Dict[str, torch.Tensor], device) -> torch.Tensor:
    """Five Conv1d(k=1) MLPs (model/chore.py:113-126) -> one fp32 buffer, k-major, first layer rows permuted/padded (csrc/query.cu)."""
    keep, wt, bt = _decoder_tables(sd)
    out = torch.empty(int(_lib.load().vt_query_wpack_floats()), dtype=torch.float32)
    _lib.call("vt_pack_weights_decoders", wt, bt, _hp(out), None)
    return out.to(device)


def pack_decoders_bwd(sd: Dict[str, torch.Tensor], device) -> torch.Tensor:
    """Transposed copies for the analytic backward pass (csrc/query.cu): per head W1b [128][640], W2b, W3b [128][128], W4b [16][128]."""
    keep, wt, bt = _decoder_tables(sd)
    out = torch.empty(int(_lib.load().vt_query_wpack_bwd_floats()), dtype=torch.float32)
    _lib.call("vt_pack_weights_decoders", wt, bt, None, _hp(out))
    return out.to(device)


def _pack_tc(sd, with_bwd: bool):
    keep, wt, _ = _decoder_tables(sd)
    h = lambda *s: torch.empty(*s, dtype=torch.float16)
    fwd = (h(5 * Q_H, Q_KB), h(5 * Q_H, Q_KB), h(2 * 5 * Q_H, Q_H), h(2 * 5 * Q_H, Q_H))
    bwd = (h(2 * 5 * Q_H, Q_H), h(2 * 5 * Q_H, Q_H), h(5 * Q_KB, Q_H), h(5 * Q_KB, Q_H)) if with_bwd else (None,) * 4
    try:
        _lib.call("vt_pack_weights_decoders_tc", wt, *(_hp(t) for t in fwd), *(None if t is None else _hp(t) for t in bwd))
    except RuntimeError as e:
        raise ValueError(str(e)) from None
    return fwd, bwd


def pack_decoders_tc(sd: Dict[str, torch.Tensor], device):
    """fp16 hi/lo planes for the tcgen05 decoder kernel: W1 [5*128, 640] (rows = head, unit; columns in the kernel's feature order) and
    W2|W3 [2*5*128, 128] (layer-major, then head) -- torch's [out][in] layout, i.e. K-major B operands."""
    fwd, _ = _pack_tc(sd, False)
    return tuple(t.to(device) for t in fwd)


def pack_decoders_tc_bwd(sd: Dict[str, torch.Tensor], device):
    """Transposed fp16 hi/lo planes for the tcgen05 backward kernel (csrc/query_bwd_tc.cu): W2^T | W3^T [2*5*128, 128] and W1^T [5*640, 128]
    (rows = feature in the kernel's order, zero rows for the padding)."""
    _, bwd = _pack_tc(sd, True)
    return tuple(t.to(device) for t in bwd)
