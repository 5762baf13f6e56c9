"""Tensor-level wrappers of the individual C-ABI operators (used by the unit tests and handy for integration work).
Every function enqueues one kernel of libvistracker_sm100a.so on the current CUDA stream; NHWC fp32 unless noted."""
from __future__ import annotations

import torch

from . import _lib
from .weights import pack_conv, pack_stem

P, S = _lib.ptr, _lib.stream_ptr


def new_stats(n, c, device):
    return torch.zeros(n, c, 2, dtype=torch.float64, device=device)


def _st(stats):
    return (P(stats), stats.stride(0) // 2) if stats is not None else (None, 0)


def stem_conv(images, weight, bias, c_off, cin, n_views, stats=None):
    B, Ctot, H, W = images.shape
    cout = weight.shape[0]
    out = torch.empty(B * n_views, H // 2, W // 2, cout, device=images.device)
    sp, sld = _st(stats)
    _lib.call("vt_stem_conv7x7s2", P(images), B, Ctot, H, W, c_off, cin, n_views, P(pack_stem(weight).to(images.device)), P(bias),
              cout, P(out), sp, sld, S())
    return out


def gn_finalize(stats, gamma, beta, count_per_channel, groups=32, eps=1e-5):
    n, C, _ = stats.shape
    ss = torch.empty(2, n, C, device=stats.device)
    _lib.call("vt_gn_finalize", P(stats), stats.stride(0) // 2, P(gamma), P(beta), n, C, groups, count_per_channel, eps, P(ss[0]),
              P(ss[1]), S())
    return ss[0], ss[1]


def affine_act(x, scale, shift, relu, stats=None):
    n, H, W, C = x.shape
    out = torch.empty(n, H, W, C, device=x.device)
    sp, sld = _st(stats)
    _lib.call("vt_affine_act", P(x), x.stride(2), P(scale), P(shift), int(relu), n, H * W, C, P(out), C, sp, sld, S())
    return out


def prep_split(x, scale, shift, relu, pad, cpad=None):
    n, H, W, C = x.shape
    cpad = cpad or (C + 63) // 64 * 64
    planes = torch.empty(2, n, H + 2 * pad, W + 2 * pad, cpad, dtype=torch.float16, device=x.device)
    ovf = torch.zeros(1, dtype=torch.int32, device=x.device)
    _lib.call("vt_prep_split", P(x), x.stride(2), P(scale), P(shift), int(relu), n, H, W, C, cpad, pad, P(planes[0]), P(planes[1]),
              P(ovf), S())
    return planes, ovf


def prep_split_gn(x, stats, gamma, beta, relu, pad, groups=32, eps=1e-5, cpad=None):
    """GroupNorm finalisation + apply + ReLU + split in one launch (vt_prep_split_gn)."""
    n, H, W, C = x.shape
    cpad = cpad or (C + 63) // 64 * 64
    planes = torch.empty(2, n, H + 2 * pad, W + 2 * pad, cpad, dtype=torch.float16, device=x.device)
    ovf = torch.zeros(1, dtype=torch.int32, device=x.device)
    _lib.call("vt_prep_split_gn", P(x), x.stride(2), P(stats), stats.stride(0) // 2, P(gamma), P(beta), groups, H * W, eps, int(relu),
              n, H, W, C, cpad, pad, P(planes[0]), P(planes[1]), P(ovf), S())
    return planes, ovf


def conv_mma(planes, H, W, pad, weight, bias=None, res=None, out=None, stats=None):
    pk = pack_conv(weight.to(planes.device))
    n = planes.shape[1]
    if out is None:
        out = torch.empty(n, H, W, pk["cout"], device=planes.device)
    sp, sld = _st(stats)
    _lib.call("vt_conv_mma", P(planes[0]), P(planes[1]), n, H, W, pk["cin_pad"], pad, pk["ks"], P(pk["hi"]), P(pk["lo"]), pk["cout"],
              P(bias), P(res), 0 if res is None else res.stride(2), P(out), out.stride(2), sp, sld, S())
    return out


def conv_ffma(x, scale, shift, relu, weight, bias=None, res=None, out=None, stats=None):
    pk = pack_conv(weight.to(x.device))
    n, H, W, C = x.shape
    if out is None:
        out = torch.empty(n, H, W, pk["cout"], device=x.device)
    sp, sld = _st(stats)
    _lib.call("vt_conv_ffma", P(x), x.stride(2), P(scale), P(shift), int(relu), n, H, W, C, pk["ks"], P(pk["ffma"]), pk["cout"],
              P(bias), P(res), 0 if res is None else res.stride(2), P(out), out.stride(2), sp, sld, S())
    return out


def add(a, b, stats=None):
    n, H, W, C = a.shape
    out = torch.empty(n, H, W, C, device=a.device)
    sp, sld = _st(stats)
    _lib.call("vt_add", P(a), a.stride(2), P(b), b.stride(2), n, H * W, C, P(out), C, sp, sld, S())
    return out


def avgpool2(x, stats=None):
    n, H, W, C = x.shape
    out = torch.empty(n, H // 2, W // 2, C, device=x.device)
    sp, sld = _st(stats)
    _lib.call("vt_avgpool2", P(x), n, H, W, C, P(out), sp, sld, S())
    return out


def upsample2x_add(low, up1, stats=None):
    n, Hl, Wl, C = low.shape
    out = torch.empty_like(up1)
    sp, sld = _st(stats)
    _lib.call("vt_upsample2x_add", P(low), P(up1), n, Hl, Wl, C, P(out), sp, sld, S())
    return out
