"""Stacked-hourglass encoder (``HGFilter``, model/HGFilters.py:56-203) as a host-side launch plan over the sm_100a
kernels of libvistracker_sm100a.so.

Layout decisions:
  * activations are NHWC fp32; every kernel that writes a tensor a GroupNorm will read also accumulates its per-channel
    (sum, sum^2) in fp64, so GroupNorm costs one tiny finalize launch instead of two extra passes over HBM;
  * GroupNorm-apply + ReLU is fused into the consumer: either into the fp16 hi/lo operand preparation of the tcgen05
    convolution, or into the tile load of the CUDA-core convolution;
  * a ConvBlock writes its three conv outputs straight into channel slices of one output tensor (no torch.cat) and the
    residual is added by the 1x1 downsample conv's epilogue or one fused add+statistics pass.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from . import _lib
from .config import EncoderDims
from .weights import MMA_KC, pack_conv, pack_stem

_I = _lib.C.c_int
GN_EPS = 1e-5
GROUPS = 32


class Act:
    """An NHWC activation [n, H, W, C] (possibly a channel slice of a wider tensor) + its statistics slot."""
    __slots__ = ("t", "stats")

    def __init__(self, t: torch.Tensor, stats: Optional[torch.Tensor]):
        self.t, self.stats = t, stats

    @property
    def n(self): return self.t.shape[0]
    @property
    def H(self): return self.t.shape[1]
    @property
    def W(self): return self.t.shape[2]
    @property
    def C(self): return self.t.shape[3]
    @property
    def ld(self): return self.t.stride(2)

    def slice(self, c0, c1):
        return Act(self.t[..., c0:c1], None if self.stats is None else self.stats[:, c0:c1])


class Operand:
    """Input of a convolution: x (raw), optional GroupNorm (the name of its affine parameters; statistics come with the Act), ReLU
    flag.  The fp16 planes for the tensor-core path are prepared once and shared by every conv that reads the same normalised
    tensor (e.g. l{i} and bl{i}); the per-(image, channel) scale/shift buffers only exist if a CUDA-core kernel asks for them."""

    def __init__(self, act: Act, bn: Optional[str] = None, relu=False):
        self.act, self.bn, self.relu = act, bn, relu
        self.scale = self.shift = None
        self.planes = {}


class StatsArena:
    """One zeroed fp64 buffer per forward pass, bump-allocated into [n, C, 2] slots."""

    def __init__(self, device, n_doubles: int):
        self.buf = torch.zeros(n_doubles, dtype=torch.float64, device=device)
        self.off = 0

    def reset(self):
        self.buf.zero_()
        self.off = 0

    def take(self, n, c):
        k = n * c * 2
        if self.off + k > self.buf.numel():
            raise RuntimeError("statistics arena exhausted")
        v = self.buf[self.off:self.off + k].view(n, c, 2)
        self.off += k
        return v


def mma_tileable(H: int, W: int) -> bool:
    """Shape rule of vt_conv_mma (include/vistracker_b200.h): 128-pixel strips of bw=min(W,128) x 128/bw."""
    bw = 128 if W >= 128 else W
    return bw >= 8 and W % bw == 0 and 128 % bw == 0 and H % (128 // bw) == 0


class HGEncoder:
    """Weights of one ``HGFilter`` re-packed for the kernels + the forward launch plan."""

    def __init__(self, sd: Dict[str, torch.Tensor], prefix: str, dims: EncoderDims, device):
        self.p, self.d, self.dev = prefix, dims, device
        self.conv: Dict[str, dict] = {}
        self.vec: Dict[str, torch.Tensor] = {}
        for k, v in sd.items():
            if not k.startswith(prefix + "."):
                continue
            name = k[len(prefix) + 1:]
            if ".downsample.0." in k:
                continue                                   # alias of bn4 (model/net_util.py:364-369)
            if v.dim() == 4 and name == "conv1.weight":
                self.stem_w = pack_stem(v).to(device)
            elif v.dim() == 4:
                # packed on the host (the checkpoint arrives there), uploaded as three plain copies: no device launches at load time
                pk = pack_conv(v.detach().cpu())
                self.conv[name[:-len(".weight")]] = {k: (t.to(device) if torch.is_tensor(t) else t) for k, t in pk.items()}
            else:
                self.vec[name] = v.float().contiguous().to(device)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=device)
        self.arena: Optional[StatsArena] = None
        self.force_ffma = os.environ.get("VT_CONV_ALGO", "") == "ffma"
        self.fuse_residual = os.environ.get("VT_FUSE_RESIDUAL", "1") != "0"
        self.launches = 0

    # ------------------------------------------------------------------ primitive launches
    def _st(self, stats):
        return (_lib.ptr(stats), stats.stride(0) // 2) if stats is not None else (None, 0)

    def _gn(self, act: Act, bn: str) -> Operand:
        """GroupNorm(32, C) of `act` with affine `bn` + ReLU, as a conv operand (not materialised)."""
        return Operand(act, bn, True)

    def _affine(self, op: Operand):
        """scale/shift buffers of a normalised operand (vt_gn_finalize) for the CUDA-core kernels; the tensor-core path derives the
        affine inside vt_prep_split_gn instead."""
        if op.bn is not None and op.scale is None:
            act = op.act
            n, C = act.n, act.C
            ss = torch.empty(2, n, C, dtype=torch.float32, device=self.dev)
            sp, ld = self._st(act.stats)
            _lib.call("vt_gn_finalize", sp, ld, _lib.ptr(self.vec[op.bn + ".weight"]), _lib.ptr(self.vec[op.bn + ".bias"]), n, C, GROUPS,
                      act.H * act.W, GN_EPS, _lib.ptr(ss[0]), _lib.ptr(ss[1]), _lib.stream_ptr())
            self.launches += 1
            op.scale, op.shift = ss[0], ss[1]
        return op.scale, op.shift

    def _conv(self, op: Operand, name: str, out: Act, bias: Optional[str] = None, res: Optional[Act] = None, stats=None,
              out2: Optional[Act] = None, res2: Optional[Act] = None):
        """out = conv(op) + bias (+ res); optionally also out2 = out + res2 (tensor-core path only), each with its statistics."""
        pk = self.conv[name]
        a = op.act
        n, H, W = a.n, a.H, a.W
        assert pk["cin"] == a.C and pk["cout"] == out.C, (name, pk["cin"], a.C, pk["cout"], out.C)
        sp, sld = self._st(stats)
        b = _lib.ptr(self.vec[bias]) if bias else None
        rp, rld = (_lib.ptr(res.t), res.ld) if res is not None else (None, 0)
        ks = pk["ks"]
        if mma_tileable(H, W) and not self.force_ffma:
            pad = ks // 2
            if pad not in op.planes:
                cpad = pk["cin_pad"]
                planes = torch.empty(2, n, H + 2 * pad, W + 2 * pad, cpad, dtype=torch.float16, device=self.dev)
                if op.bn is not None:
                    sp_in, ld_in = self._st(a.stats)
                    _lib.call("vt_prep_split_gn", _lib.ptr(a.t), a.ld, sp_in, ld_in, _lib.ptr(self.vec[op.bn + ".weight"]),
                              _lib.ptr(self.vec[op.bn + ".bias"]), GROUPS, H * W, GN_EPS, int(op.relu), n, H, W, a.C, cpad, pad,
                              _lib.ptr(planes[0]), _lib.ptr(planes[1]), _lib.ptr(self.overflow), _lib.stream_ptr())
                else:
                    sc, sh = self._affine(op)
                    _lib.call("vt_prep_split", _lib.ptr(a.t), a.ld, _lib.ptr(sc), _lib.ptr(sh), int(op.relu), n, H, W,
                              a.C, cpad, pad, _lib.ptr(planes[0]), _lib.ptr(planes[1]), _lib.ptr(self.overflow), _lib.stream_ptr())
                self.launches += 1
                op.planes[pad] = planes
            planes = op.planes[pad]
            if out2 is None:
                _lib.call("vt_conv_mma", _lib.ptr(planes[0]), _lib.ptr(planes[1]), n, H, W, pk["cin_pad"], pad, ks,
                          _lib.ptr(pk["hi"]), _lib.ptr(pk["lo"]), pk["cout"], b, rp, rld, _lib.ptr(out.t), out.ld, sp, sld,
                          _lib.stream_ptr())
            else:
                sp2, sld2 = self._st(out2.stats)
                _lib.call("vt_conv_mma_dual", _lib.ptr(planes[0]), _lib.ptr(planes[1]), n, H, W, pk["cin_pad"], pad, ks,
                          _lib.ptr(pk["hi"]), _lib.ptr(pk["lo"]), pk["cout"], b, rp, rld, _lib.ptr(out.t), out.ld, sp, sld,
                          _lib.ptr(out2.t), out2.ld, _lib.ptr(res2.t), res2.ld, sp2, sld2, _lib.stream_ptr())
        else:
            assert out2 is None, "the fused second output exists on the tensor-core path only"
            sc, sh = self._affine(op)
            _lib.call("vt_conv_ffma", _lib.ptr(a.t), a.ld, _lib.ptr(sc), _lib.ptr(sh), int(op.relu), n, H, W, a.C,
                      ks, _lib.ptr(pk["ffma"]), pk["cout"], b, rp, rld, _lib.ptr(out.t), out.ld, sp, sld, _lib.stream_ptr())
        self.launches += 1

    def _new(self, n, H, W, C, with_stats=True) -> Act:
        t = torch.empty(n, H, W, C, dtype=torch.float32, device=self.dev)
        return Act(t, self.arena.take(n, C) if with_stats else None)

    # ------------------------------------------------------------------ network blocks
    def conv_block(self, x: Act, name: str, cout: int) -> Act:
        """ConvBlock.forward, model/net_util.py:374-396."""
        n, H, W = x.n, x.H, x.W
        half, quarter = cout // 2, cout // 4
        out = self._new(n, H, W, cout)                      # out.stats: statistics of the block OUTPUT (after residual)
        raw_stats = self.arena.take(n, cout)                # statistics of the raw conv1 / conv2 outputs (bn2 / bn3 inputs)
        if self.fuse_residual and f"{name}.downsample.2" not in self.conv and mma_tileable(H, W) and not self.force_ffma:
            # identity residual, tensor-core path: every conv epilogue also writes its slice of (cat + x) -- no add pass.
            # Default since the persistent conv kernel overlaps its epilogue with the next tile's mainloop (2.2 ms / step faster on
            # B200; with the one-tile-per-CTA kernel it was 3 ms slower, profiles/r01e_*).  VT_FUSE_RESIDUAL=0 restores vt_add.
            raw = torch.empty(n, H, W, half + quarter, dtype=torch.float32, device=self.dev)
            r1 = Act(raw[..., :half], raw_stats[:, :half])
            r2 = Act(raw[..., half:], raw_stats[:, half:half + quarter])
            self._conv(self._gn(x, f"{name}.bn1"), f"{name}.conv1", r1, stats=r1.stats, out2=out.slice(0, half), res2=x.slice(0, half))
            self._conv(self._gn(r1, f"{name}.bn2"), f"{name}.conv2", r2, stats=r2.stats, out2=out.slice(half, half + quarter),
                       res2=x.slice(half, half + quarter))
            o3 = out.slice(half + quarter, cout)
            self._conv(self._gn(r2, f"{name}.bn3"), f"{name}.conv3", o3, res=x.slice(half + quarter, cout), stats=o3.stats)
            return out
        o1 = Act(out.t[..., :half], raw_stats[:, :half])
        o2 = Act(out.t[..., half:half + quarter], raw_stats[:, half:half + quarter])
        o3 = Act(out.t[..., half + quarter:], None)
        self._conv(self._gn(x, f"{name}.bn1"), f"{name}.conv1", o1, stats=o1.stats)
        self._conv(self._gn(o1, f"{name}.bn2"), f"{name}.conv2", o2, stats=o2.stats)
        self._conv(self._gn(o2, f"{name}.bn3"), f"{name}.conv3", o3)
        if f"{name}.downsample.2" in self.conv:
            self._conv(self._gn(x, f"{name}.bn4"), f"{name}.downsample.2", out, res=out, stats=out.stats)
        else:
            sp, sld = self._st(out.stats)
            _lib.call("vt_add", _lib.ptr(out.t), out.ld, _lib.ptr(x.t), x.ld, n, H * W, cout, _lib.ptr(out.t), out.ld, sp, sld,
                      _lib.stream_ptr())
            self.launches += 1
        return out

    def pool(self, x: Act) -> Act:
        out = self._new(x.n, x.H // 2, x.W // 2, x.C)
        assert x.t.is_contiguous()
        sp, sld = self._st(out.stats)
        _lib.call("vt_avgpool2", _lib.ptr(x.t), x.n, x.H, x.W, x.C, _lib.ptr(out.t), sp, sld, _lib.stream_ptr())
        self.launches += 1
        return out

    def hourglass(self, x: Act, name: str, level: int) -> Act:
        """HourGlass._forward, model/HGFilters.py:26-50."""
        C = x.C
        up1 = self.conv_block(x, f"{name}.b1_{level}", C)
        low = self.conv_block(self.pool(x), f"{name}.b2_{level}", C)
        low = self.hourglass(low, name, level - 1) if level > 1 else self.conv_block(low, f"{name}.b2_plus_{level}", C)
        low = self.conv_block(low, f"{name}.b3_{level}", C)
        out = self._new(x.n, x.H, x.W, C)
        sp, sld = self._st(out.stats)
        _lib.call("vt_upsample2x_add", _lib.ptr(low.t), _lib.ptr(up1.t), x.n, low.H, low.W, C, _lib.ptr(out.t), sp, sld,
                  _lib.stream_ptr())
        self.launches += 1
        return out

    def forward(self, images: torch.Tensor, c_off: int, n_views: int):
        """HGFilter.forward (model/HGFilters.py:162-203) in eval mode.  images: NCHW [B, Ctot, H, W] fp32 on the device.
        Encoder image n = view*B + b reads channels c_off + view*in_ch + [0, in_ch).  Returns (last-stack features
        [n, H/4, W/4, out_ch], tmpx [n, H/2, W/2, stem_ch]) as NHWC tensors."""
        d = self.d
        B, Ctot, H, W = images.shape
        if H % 16 or W % 16:
            raise ValueError(f"frame size {H}x{W} must be a multiple of 16 (two hourglass levels below H/4)")
        n = B * n_views
        need = n * 2 * (30 * d.num_stack + 16) * d.feat_ch                # >= 2 slots per ConvBlock + pools/upsamples/heads
        if self.arena is None or self.arena.buf.numel() < need:
            self.arena = StatsArena(self.dev, need)
        self.arena.reset()
        self.launches = 1
        images = images.contiguous()
        raw0 = self._new(n, H // 2, W // 2, d.stem_ch)
        sp, sld = self._st(raw0.stats)
        _lib.call("vt_stem_conv7x7s2", _lib.ptr(images), B, Ctot, H, W, c_off, d.in_ch, n_views, _lib.ptr(self.stem_w),
                  _lib.ptr(self.vec["conv1.bias"]), d.stem_ch, _lib.ptr(raw0.t), sp, sld, _lib.stream_ptr())
        g_sc, g_sh = self._affine(self._gn(raw0, "bn1"))
        tmpx = self._new(n, H // 2, W // 2, d.stem_ch)
        sp, sld = self._st(tmpx.stats)
        _lib.call("vt_affine_act", _lib.ptr(raw0.t), raw0.ld, _lib.ptr(g_sc), _lib.ptr(g_sh), 1, n, raw0.H * raw0.W,
                  d.stem_ch, _lib.ptr(tmpx.t), tmpx.ld, sp, sld, _lib.stream_ptr())
        self.launches += 2
        x = self.pool(self.conv_block(tmpx, "conv2", 128))
        x = self.conv_block(self.conv_block(x, "conv3", 128), "conv4", d.feat_ch)
        previous = x
        out = None
        for i in range(d.num_stack):
            ll = self.conv_block(self.hourglass(previous, f"m{i}", d.depth), f"top_m_{i}", d.feat_ch)
            cl = self._new(n, ll.H, ll.W, d.feat_ch)
            self._conv(Operand(ll), f"conv_last{i}", cl, bias=f"conv_last{i}.bias", stats=cl.stats)
            llo = self._gn(cl, f"bn_end{i}")                                  # relu(bn_end(conv_last(ll))), shared by l / bl
            out = self._new(n, ll.H, ll.W, d.out_ch, with_stats=False)
            self._conv(llo, f"l{i}", out, bias=f"l{i}.bias")
            if i < d.num_stack - 1:
                nxt = self._new(n, ll.H, ll.W, d.feat_ch)
                self._conv(llo, f"bl{i}", nxt, bias=f"bl{i}.bias", res=previous)
                self._conv(Operand(out), f"al{i}", nxt, bias=f"al{i}.bias", res=nxt, stats=nxt.stats)
                previous = nxt
        return out.t, tmpx.t

    def check_overflow(self):
        if int(self.overflow.item()) != 0:
            self.overflow.zero_()
            raise RuntimeError("an activation exceeded the fp16 range in the fp16x2 tensor-core convolution path")
