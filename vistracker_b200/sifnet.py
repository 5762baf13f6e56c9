"""Drop-in for the reference's SIF-Net object on B200.

``CHORETriplaneVisibility`` below keeps the public surface of ``model.CHORETriplaneVisibility``
(model/chore_tri_vis.py:16, model/chore_triplane.py:16-251, model/chore.py:17-223, model/BasePIFuNet.py:6-70):
constructor ``(opt, projection_mode, error_term, rank, num_parts, hidden_dim)``, ``load_state_dict`` with the
reference's 706 checkpoint keys, ``filter(images)``, ``query(points, crop_center=None, body_center=...)``,
``query_features``, ``get_preds``, ``project_points``, ``triplane_project``, ``OUT_DIST`` and the cached-map attributes.
All arithmetic runs in the sm_100a kernels of libvistracker_sm100a.so; there is no PyTorch fallback.
"""
from __future__ import annotations

import ctypes
from collections import OrderedDict
from typing import Dict, Optional

import torch

from . import _lib
from .config import SIFNetDims, resolve_dims
from .encoder import HGEncoder
from .synth import sifnet_spec
from .weights import pack_decoders, pack_decoders_bwd, pack_decoders_tc, pack_decoders_tc_bwd

N_OUT = 29          # df 2 | pca 9 | parts 14 | centers 3 | visibility 1


HEAD_SLICES = ((0, 2), (2, 11), (11, 25), (25, 28), (28, 29))      # df | pca | parts | centers | visibility in the packed output


class _QueryFn(torch.autograd.Function):
    """query() is differentiable w.r.t. the points only (network weights are frozen at inference,
    recon/gen/generator.py:53-54).  The five heads are separate outputs so that autograd tells us which ones carry a cotangent:
    the backward kernel recomputes and back-propagates only those."""

    @staticmethod
    def forward(ctx, net: "CHORETriplaneVisibility", points, crop_center, body_center):
        out, xy = net._query_raw(points, crop_center, body_center, want_xy=True)
        ctx.net = net
        ctx.maps = net._maps                       # the backward differentiates against the maps of THIS forward, not of a later filter()
        ctx.save_for_backward(points, crop_center, body_center)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(xy)
        return tuple(out[:, lo:hi] for lo, hi in HEAD_SLICES) + (xy,)

    @staticmethod
    def backward(ctx, *grads):
        points, crop_center, body_center = ctx.saved_tensors
        mask = sum(1 << h for h in range(5) if grads[h] is not None)
        if mask == 0:
            return None, torch.zeros_like(points), None, None
        B, N = points.shape[0], points.shape[1]
        g_out = torch.zeros(B, N_OUT, N, dtype=torch.float32, device=points.device)
        for h, (lo, hi) in enumerate(HEAD_SLICES):
            if grads[h] is not None:
                g_out[:, lo:hi] = grads[h]
        g_pts = ctx.net._query_backward(points, crop_center, body_center, g_out, head_mask=mask, maps=ctx.maps)
        return None, g_pts, None, None


class _QueryLossFn(torch.autograd.Function):
    """Per-point fitting-loss terms with their point gradients from ONE launch (vt_query_losses_tc); backward is two multiplies."""

    @staticmethod
    def forward(ctx, net: "CHORETriplaneVisibility", points, crop_center, body_center, df_channel, clamp_max, part_labels, fwd_mask):
        vals_df, g_df, vals_ce, g_ce, out_fwd = net._query_losses_raw(points, crop_center, body_center, df_channel, clamp_max, part_labels,
                                                                      fwd_mask)
        ctx.has_ce = vals_ce is not None
        ctx.save_for_backward(*((g_df, g_ce) if ctx.has_ce else (g_df,)))
        if not ctx.has_ce:
            vals_ce = vals_df.new_zeros(())
            ctx.mark_non_differentiable(vals_ce)
        if out_fwd is None:
            out_fwd = vals_df.new_zeros(())
        ctx.mark_non_differentiable(out_fwd)
        return vals_df, vals_ce, out_fwd

    @staticmethod
    def backward(ctx, grad_df, grad_ce, _grad_out):
        saved = ctx.saved_tensors
        g = grad_df.unsqueeze(-1) * saved[0]
        if ctx.has_ce and grad_ce is not None:
            g = g + grad_ce.unsqueeze(-1) * saved[1]
        return None, g, None, None, None, None, None, None


class CHORETriplaneVisibility:
    """B200-native SIF-Net (tri-vis).  Not an nn.Module: parameters live as kernel-layout device buffers."""

    def __init__(self, opt, projection_mode="perspective", error_term=None, rank=-1, num_parts=14, hidden_dim=128,
                 device: Optional[torch.device] = None):
        if projection_mode != "perspective":
            raise NotImplementedError("only the perspective camera of the tri-vis model is built")
        self.opt = opt
        self.name = "chore"
        self.dims: SIFNetDims = resolve_dims(opt, num_parts)
        if self.dims.hidden != 128 or self.dims.feature_size != 611:
            raise NotImplementedError("the fused decoder kernel is built for hidden_dim=128 and the 611-d tri-vis feature")
        self.device = torch.device(device) if device is not None else torch.device("cuda", int(getattr(opt, "gpu_id", 0)))
        self.feature_size, self.hidden_dim, self.z_feat = self.dims.feature_size, self.dims.hidden, opt.z_feat
        self.shared_encoder = True
        self.OUT_DIST = self.dims.out_dist
        self.training = False
        self._expected = OrderedDict((k, tuple(s)) for k, s, _ in sifnet_spec(self.dims))
        self._rgb: Optional[HGEncoder] = None
        self._tri: Optional[HGEncoder] = None
        self._wpack: Optional[torch.Tensor] = None
        d = self.dims
        self._cam7 = (ctypes.c_float * 7)(d.fx_px, d.fy_px, d.cx_px, d.cy_px, d.crop_size, d.z0, d.out_dist)
        # buffers mirrored from the reference object
        self._maps = None
        self.preds = None
        self.intermediate_preds_list = []
        self.points = self.points_xy = self.crop_center = self.local_feat_list = self.input_images = None
        self.defer_checks = False
        import os
        self.query_on_cuda_cores = os.environ.get("VT_QUERY", "") == "ffma"      # cross-check switch; default = tensor cores
        self.filter_streams = int(os.environ.get("VT_FILTER_STREAMS", "1"))   # 2 = encoders on two streams (measured slower: static tile ranges)
        self._side_stream = None
        self.use_graph = os.environ.get("VT_FILTER_GRAPH", "1") != "0"
        self._graphs = {}

    # ------------------------------------------------------------------ nn.Module-like surface
    def to(self, device):
        self.device = torch.device(device)
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", 0 if device is None else device))

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        if mode:
            raise NotImplementedError("training is outside the B200 hot path (SURVEY.md section 2 row 13)")
        return self

    def parameters(self):
        return iter(())            # weights are frozen kernel buffers; nothing for an optimiser to touch

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True):
        """Accepts the reference checkpoint's ``model_state_dict`` (keys as in tests/golden/sifnet_keys.json; a leading
        ``module.`` from DataParallel training is stripped like recon/gen/generator.py:296-308 does)."""
        sd = OrderedDict((k[7:] if k.startswith("module.") else k, v) for k, v in sd.items())
        missing = [k for k in self._expected if k not in sd]
        unexpected = [k for k in sd if k not in self._expected]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict: missing {missing[:5]} unexpected {unexpected[:5]}")
        for k, shape in self._expected.items():
            if k in sd and tuple(sd[k].shape) != shape:
                raise RuntimeError(f"size mismatch for {k}: checkpoint {tuple(sd[k].shape)} vs model {shape}")
        if self.device.type != "cuda":
            raise RuntimeError("vistracker_b200 has no CPU path: move the model to a CUDA device before loading weights")
        _lib.load()
        with torch.cuda.device(self.device):
            self._rgb = HGEncoder(sd, "image_filter", self.dims.rgb, self.device)
            self._tri = HGEncoder(sd, "triplane_encoder", self.dims.tri, self.device)
            self._wpack = pack_decoders(sd, self.device)
            self._wpack_bwd = pack_decoders_bwd(sd, self.device)
            self._wtc = pack_decoders_tc(sd, self.device)
            self._wtc_bwd = pack_decoders_tc_bwd(sd, self.device)
            self._q_overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
        assert self._wpack.numel() == _lib.load().vt_query_wpack_floats()
        assert self._wpack_bwd.numel() == _lib.load().vt_query_wpack_bwd_floats()
        return self

    # ------------------------------------------------------------------ filter
    def filter(self, images: torch.Tensor, into=None):
        """CHORETriplane.filter (model/chore_triplane.py:60-95): images [B, 8, H, W] = RGB, person mask, object mask,
        3 triplane renderings.  The three triplane views go through the shared encoder as one batch of 3B.

        The ~440 kernel launches of the two encoders are captured once per input shape in a CUDA graph and replayed (static input
        buffer, graph-private activations, outputs copied out): same kernels, same results, no host launch cost and ~1 us instead
        of ~3 us between dependent kernels.  VT_FILTER_GRAPH=0 launches eagerly."""
        assert images.shape[1] == 8, f"given image shape invalide: {images.shape}"
        if self._rgb is None:
            raise RuntimeError("load_state_dict() must be called before filter()")
        images = images.to(self.device, torch.float32).contiguous()
        with torch.cuda.device(self.device):
            if self.use_graph and self.filter_streams == 1 and not torch.cuda.is_current_stream_capturing():
                key = tuple(images.shape)
                entry = self._graphs.get(key)
                if entry is None:
                    static_in = torch.empty_like(images)
                    static_in.copy_(images)
                    # every captured graph owns its GroupNorm-statistics arenas: a later, larger input shape must not replace (and free) a buffer
                    # whose address an earlier graph still memsets and accumulates into on replay
                    for e in (self._rgb, self._tri):
                        e.arena = None
                    self._filter_eager(static_in)                       # warm-up: lazy one-time initialisation stays outside the capture
                    torch.cuda.current_stream().synchronize()
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        maps = self._filter_eager(static_in)
                    entry = (graph, static_in, maps, self._rgb.launches + self._tri.launches, (self._rgb.arena, self._tri.arena))
                    for e in (self._rgb, self._tri):
                        e.arena = None                                  # eager calls after this allocate their own
                    if len(self._graphs) >= 4:                          # bound the memory held by graph-private pools
                        self._graphs.pop(next(iter(self._graphs)))
                    self._graphs[key] = entry
                graph, static_in, static_maps, launches, _arenas = entry
                static_in.copy_(images)
                graph.replay()
                # the graph writes into its private buffers; hand out copies so that maps kept from an earlier filter() call stay valid,
                # as with the reference's freshly allocated tensors (0.57 GB at B = 8: ~0.2 ms, < 1 % of the step)
                if into is None:
                    maps = tuple(t.clone() for t in static_maps)
                else:
                    # chunked filter of a larger batch (recon_driver.filter_batch): this chunk's maps go straight into their rows of the
                    # whole-batch tensors -- one copy instead of clone + torch.cat
                    full, s0, Bt = into
                    n = images.shape[0]
                    full[0][s0:s0 + n].copy_(static_maps[0]); full[1][s0:s0 + n].copy_(static_maps[1])
                    for v in range(3):
                        full[2][v * Bt + s0:v * Bt + s0 + n].copy_(static_maps[2][v * n:(v + 1) * n])
                        full[3][v * Bt + s0:v * Bt + s0 + n].copy_(static_maps[3][v * n:(v + 1) * n])
                    maps = static_maps
                self.launches_filter = launches + 4
            else:
                maps = self._filter_eager(images)
                self.launches_filter = self._rgb.launches + self._tri.launches
            if not self.defer_checks:
                self.check()
        self.input_images = images
        self._maps = maps

    def _filter_eager(self, images: torch.Tensor):
        """One pass of both encoders on the current stream; returns (im_feat, tmpx, tri_tmpx, tri_feat) as NHWC tensors."""
        if self.filter_streams >= 4:
            # experiment: RGB encoder + the three triplane views as four independent n = B chains on four streams, so that the
            # HBM-bound passes of one chain overlap the tensor-bound convolutions of the others
            import copy
            main = torch.cuda.current_stream()
            if self._side_stream is None:
                self._side_stream = [torch.cuda.Stream(device=self.device) for _ in range(3)]
                self._tri_views = []
                for _ in range(3):
                    e = copy.copy(self._tri)
                    e.arena, e.overflow = None, torch.zeros(1, dtype=torch.int32, device=self.device)
                    self._tri_views.append(e)
            outs = []
            for v, (st, enc) in enumerate(zip(self._side_stream, self._tri_views)):
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    outs.append(enc.forward(images, 5 + v, 1))
            im_feat, tmpx = self._rgb.forward(images, 0, 1)
            for st in self._side_stream:
                main.wait_stream(st)
            for o in outs:                              # allocated on a side stream, consumed on the caller's stream
                o[0].record_stream(main); o[1].record_stream(main)
            tri_feat = torch.cat([o[0] for o in outs], 0)
            tri_tmpx = torch.cat([o[1] for o in outs], 0)
            self._tri.launches = sum(e.launches for e in self._tri_views)
        elif self.filter_streams > 1:
            # the RGB encoder (n = B) and the shared triplane encoder (n = 3B) are independent until query(): run them on two
            # streams so the small-map layers of one fill the SMs the other leaves idle and HBM-bound passes overlap MMA-bound ones
            main = torch.cuda.current_stream()
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream(device=self.device)
            side = self._side_stream
            side.wait_stream(main)
            with torch.cuda.stream(side):
                im_feat, tmpx = self._rgb.forward(images, 0, 1)
            tri_feat, tri_tmpx = self._tri.forward(images, 5, 3)
            main.wait_stream(side)
        else:
            im_feat, tmpx = self._rgb.forward(images, 0, 1)
            tri_feat, tri_tmpx = self._tri.forward(images, 5, 3)
        return im_feat, tmpx, tri_tmpx, tri_feat

    def check(self):
        self._rgb.check_overflow()
        self._tri.check_overflow()
        for e in getattr(self, "_tri_views", []):
            e.check_overflow()
        if int(self._q_overflow.item()) != 0:
            self._q_overflow.zero_()
            raise RuntimeError("a feature / activation exceeded the fp16 range in the tensor-core decoder path")

    # reference attribute names, as NCHW-shaped views of the NHWC buffers (no copy)
    @property
    def im_feat_list(self):
        return None if self._maps is None else [self._maps[0].permute(0, 3, 1, 2)]

    @property
    def tmpx(self):
        return None if self._maps is None else self._maps[1].permute(0, 3, 1, 2)

    @property
    def triplane_tmpx(self):
        if self._maps is None:
            return None
        B = self._maps[0].shape[0]
        return [self._maps[2][v * B:(v + 1) * B].permute(0, 3, 1, 2) for v in range(3)]

    @property
    def triplane_feat_list(self):
        if self._maps is None:
            return None
        B = self._maps[0].shape[0]
        return [[self._maps[3][v * B:(v + 1) * B].permute(0, 3, 1, 2)] for v in range(3)]

    def get_im_feat(self):
        return self.im_feat_list[-1]

    # ------------------------------------------------------------------ query
    def _query_raw(self, points, crop_center, body_center, want_xy=False, want_feat=False):
        if self._maps is None:
            raise RuntimeError("filter() must be called before query()")
        im_feat, tmpx, tri_tmpx, tri_feat = self._maps
        B, N = points.shape[0], points.shape[1]
        if B != im_feat.shape[0]:
            raise ValueError(f"points batch {B} != filtered batch {im_feat.shape[0]}")
        pts = points.detach().to(self.device, torch.float32).contiguous()
        cc = crop_center.to(self.device, torch.float32).contiguous()
        bc = body_center.to(self.device, torch.float32).contiguous()
        out = torch.empty(B, N_OUT, N, dtype=torch.float32, device=self.device)
        xy = torch.empty(B, 2, N, dtype=torch.float32, device=self.device) if want_xy else None
        feat = torch.empty(B, self.feature_size, N, dtype=torch.float32, device=self.device) if want_feat else None
        d = self.dims
        if not want_feat and not self.query_on_cuda_cores:
            # decoders on the tensor cores (csrc/query_tc.cu); the feature dump only exists in the CUDA-core kernel
            with torch.cuda.device(self.device):
                _lib.call("vt_query_fwd_tc", _lib.ptr(pts), _lib.ptr(cc), _lib.ptr(bc), B, N, _lib.ptr(im_feat), _lib.ptr(tmpx),
                          _lib.ptr(tri_tmpx), _lib.ptr(tri_feat), im_feat.shape[1], im_feat.shape[2], tmpx.shape[1], tmpx.shape[2],
                          self._cam7, _lib.ptr(self._wpack), *(_lib.ptr(t) for t in self._wtc), _lib.ptr(out), _lib.ptr(xy),
                          _lib.ptr(self._q_overflow), _lib.stream_ptr())
            return out, xy
        with torch.cuda.device(self.device):
            _lib.call("vt_query_fwd", _lib.ptr(pts), _lib.ptr(cc), _lib.ptr(bc), B, N, _lib.ptr(im_feat), _lib.ptr(tmpx),
                      _lib.ptr(tri_tmpx), _lib.ptr(tri_feat), im_feat.shape[1], im_feat.shape[2], tmpx.shape[1], tmpx.shape[2],
                      d.rgb.out_ch, d.rgb.stem_ch, d.tri.stem_ch, d.tri.out_ch, self._cam7, _lib.ptr(self._wpack), _lib.ptr(out),
                      _lib.ptr(feat), _lib.ptr(xy), _lib.stream_ptr())
        if want_feat:
            return feat, xy
        return out, xy

    def _query_backward(self, points, crop_center, body_center, g_out, head_mask: int = 31, maps=None):
        """d(sum g_out * out)/d(points) with the maps of the last filter() call (csrc/query.cu: query_bwd_kernel), restricted to
        the heads in ``head_mask`` (bit h set = head h has a non-zero cotangent).  ``maps``: the feature maps of the forward pass (autograd
        keeps them; default = the current ones)."""
        im_feat, tmpx, tri_tmpx, tri_feat = self._maps if maps is None else maps
        B, N = points.shape[0], points.shape[1]
        pts = points.detach().to(self.device, torch.float32).contiguous()
        cc = crop_center.to(self.device, torch.float32).contiguous()
        bc = body_center.to(self.device, torch.float32).contiguous()
        g = g_out.to(self.device, torch.float32).contiguous()
        g_pts = torch.empty(B, N, 3, dtype=torch.float32, device=self.device)
        d = self.dims
        with torch.cuda.device(self.device):
            if not self.query_on_cuda_cores:
                _lib.call("vt_query_bwd_tc", _lib.ptr(pts), _lib.ptr(cc), _lib.ptr(bc), B, N, _lib.ptr(im_feat), _lib.ptr(tmpx),
                          _lib.ptr(tri_tmpx), _lib.ptr(tri_feat), im_feat.shape[1], im_feat.shape[2], tmpx.shape[1], tmpx.shape[2],
                          self._cam7, _lib.ptr(self._wpack), *(_lib.ptr(t) for t in self._wtc), *(_lib.ptr(t) for t in self._wtc_bwd),
                          _lib.ptr(g), int(head_mask), _lib.ptr(g_pts), _lib.ptr(self._q_overflow), _lib.stream_ptr())
                return g_pts
            _lib.call("vt_query_bwd_heads", _lib.ptr(pts), _lib.ptr(cc), _lib.ptr(bc), B, N, _lib.ptr(im_feat), _lib.ptr(tmpx),
                      _lib.ptr(tri_tmpx), _lib.ptr(tri_feat), im_feat.shape[1], im_feat.shape[2], tmpx.shape[1], tmpx.shape[2],
                      d.rgb.out_ch, d.rgb.stem_ch, d.tri.stem_ch, d.tri.out_ch, self._cam7, _lib.ptr(self._wpack),
                      _lib.ptr(self._wpack_bwd), _lib.ptr(g), int(head_mask), _lib.ptr(g_pts), _lib.stream_ptr())
        return g_pts

    def _query_losses_raw(self, points, crop_center, body_center, df_channel, clamp_max, part_labels, fwd_mask=0):
        im_feat, tmpx, tri_tmpx, tri_feat = self._maps
        B, N = points.shape[0], points.shape[1]
        pts = points.detach().to(self.device, torch.float32).contiguous()
        cc = crop_center.to(self.device, torch.float32).contiguous()
        bc = body_center.to(self.device, torch.float32).contiguous()
        vals_df = torch.empty(B, N, dtype=torch.float32, device=self.device)
        g_df = torch.empty(B, N, 3, dtype=torch.float32, device=self.device)
        labels = vals_ce = g_ce = None
        if part_labels is not None:
            labels = part_labels.to(self.device, torch.int64).contiguous()
            if tuple(labels.shape) != (B, N):
                raise ValueError(f"part_labels must be [B, N] = {(B, N)}, got {tuple(labels.shape)}")
            vals_ce, g_ce = torch.empty_like(vals_df), torch.empty_like(g_df)
        out_fwd = torch.empty(B, N_OUT, N, dtype=torch.float32, device=self.device) if fwd_mask else None
        self.enqueue_query_losses(pts, cc, bc, df_channel, clamp_max, labels, vals_df, g_df, vals_ce, g_ce, fwd_mask, out_fwd)
        return vals_df, g_df, vals_ce, g_ce, out_fwd

    def enqueue_query_losses(self, pts, cc, bc, df_channel, clamp_max, labels, vals_df, g_df, vals_ce, g_ce, fwd_mask=0, out_fwd=None, maps=None):
        """The bare ``vt_query_losses_tc`` launch on caller-owned, contiguous fp32 device buffers (pts [B,N,3], cc [B,2], bc [B,3], labels int64
        [B,N] or None, outputs as in ``_query_losses_raw``): no allocation, no host synchronisation -- the form the CUDA-graph optimisation
        steps (recon_steps.py) capture.  ``maps`` defaults to the maps of the last ``filter()`` call."""
        im_feat, tmpx, tri_tmpx, tri_feat = self._maps if maps is None else maps
        B, N = pts.shape[0], pts.shape[1]
        if B != im_feat.shape[0]:
            raise ValueError(f"points batch {B} != filtered batch {im_feat.shape[0]}")
        with torch.cuda.device(self.device):
            _lib.call("vt_query_losses_tc", _lib.ptr(pts), _lib.ptr(cc), _lib.ptr(bc), B, N, _lib.ptr(im_feat), _lib.ptr(tmpx),
                      _lib.ptr(tri_tmpx), _lib.ptr(tri_feat), im_feat.shape[1], im_feat.shape[2], tmpx.shape[1], tmpx.shape[2],
                      self._cam7, _lib.ptr(self._wpack), *(_lib.ptr(t) for t in self._wtc), *(_lib.ptr(t) for t in self._wtc_bwd),
                      int(df_channel), float(clamp_max), _lib.ptr(labels), _lib.ptr(vals_df), _lib.ptr(g_df), _lib.ptr(vals_ce),
                      _lib.ptr(g_ce), int(fwd_mask), _lib.ptr(out_fwd), _lib.ptr(self._q_overflow), _lib.stream_ptr())

    def enqueue_query_losses_merged(self, pts, cc, bc, df_channel, clamp_max, labels, w_df_ptr, w_df_mul, w_ce_ptr, w_ce_mul, vals_df, vals_ce, g_points,
                                    maps=None):
        """``vt_query_losses_merged_tc``: both loss heads in one pass with ONE backward gather; ``w_*_ptr`` are raw device addresses of fp32
        scalars (e.g. words of the optimisation step's control block), ``w_*_mul`` host factors; g_points = w_df d clamp(df) + w_ce d CE."""
        im_feat, tmpx, tri_tmpx, tri_feat = self._maps if maps is None else maps
        B, N = pts.shape[0], pts.shape[1]
        if B != im_feat.shape[0]:
            raise ValueError(f"points batch {B} != filtered batch {im_feat.shape[0]}")
        with torch.cuda.device(self.device):
            _lib.call("vt_query_losses_merged_tc", _lib.ptr(pts), _lib.ptr(cc), _lib.ptr(bc), B, N, _lib.ptr(im_feat), _lib.ptr(tmpx),
                      _lib.ptr(tri_tmpx), _lib.ptr(tri_feat), im_feat.shape[1], im_feat.shape[2], tmpx.shape[1], tmpx.shape[2],
                      self._cam7, _lib.ptr(self._wpack), *(_lib.ptr(t) for t in self._wtc), *(_lib.ptr(t) for t in self._wtc_bwd),
                      int(df_channel), float(clamp_max), _lib.ptr(labels), ctypes.c_void_p(w_df_ptr), float(w_df_mul), ctypes.c_void_p(w_ce_ptr),
                      float(w_ce_mul), _lib.ptr(vals_df), _lib.ptr(vals_ce), _lib.ptr(g_points), _lib.ptr(self._q_overflow), _lib.stream_ptr())

    def query_losses(self, points, crop_center=None, df_channel=0, clamp_max=0.1, part_labels=None, also=(), **kwargs):
        """The query-dependent loss terms of the fitters as per-point tensors, differentiable w.r.t. ``points``:
        ``clamp(df[:, df_channel], max=clamp_max)`` [B, N] and (with ``part_labels`` [B, N]) ``F.cross_entropy(parts, labels,
        reduction='none')`` [B, N] -- the quantities of recon_fit_behave.py:471-476 / recon_fit_trivis_full.py:235, computed with their
        gradients by one fused launch instead of query() + autograd (an extension of the reference interface; results match
        ``query()`` followed by the same torch expressions).  ``also``: names of further heads ('pca', 'parts', 'centers', 'visibility')
        whose predictions are wanted without gradient; they ride on the same launch and come back as a third value, a dict."""
        body_center = kwargs.get("body_center")
        if crop_center is None or body_center is None:
            raise ValueError("query_losses() needs crop_center and body_center")
        if self._maps is None:
            raise RuntimeError("filter() must be called before query_losses()")
        names = ("df", "pca", "parts", "centers", "visibility")
        if self.query_on_cuda_cores:             # cross-check path: compose the same terms from query() + autograd
            self.query(points, crop_center=crop_center, body_center=body_center)
            df, _, parts = self.preds[0], self.preds[1], self.preds[2]
            vals_df = torch.clamp(df[:, df_channel, :], max=clamp_max)
            vals_ce = None
            if part_labels is not None:
                vals_ce = torch.nn.functional.cross_entropy(parts, part_labels.to(parts.device), reduction="none")
            if also:
                full = dict(zip(names, self.preds))
                return vals_df, vals_ce, {h: full[h].detach() for h in also}
            return vals_df, vals_ce
        fwd_mask = sum(1 << names.index(h) for h in also)
        vals_df, vals_ce, out_fwd = _QueryLossFn.apply(self, points, crop_center, body_center, df_channel, clamp_max, part_labels, fwd_mask)
        vals_ce = vals_ce if part_labels is not None else None
        if also:
            B, N = points.shape[0], points.shape[1]
            extra = {}
            for h in also:
                lo, hi = HEAD_SLICES[names.index(h)]
                extra[h] = out_fwd[:, lo:hi].reshape(B, 3, 3, N) if h == "pca" else out_fwd[:, lo:hi]
            return vals_df, vals_ce, extra
        return vals_df, vals_ce

    def query_heads(self, points, heads, crop_center=None, **kwargs):
        """Forward-only query of a subset of the decoder heads (names from 'df', 'pca', 'parts', 'centers', 'visibility'): returns a
        dict name -> tensor shaped like the entries of get_preds().  One gather pass serves up to four heads."""
        names = ("df", "pca", "parts", "centers", "visibility")
        mask = sum(1 << names.index(h) for h in heads)
        body_center = kwargs.get("body_center")
        im_feat, tmpx, tri_tmpx, tri_feat = self._maps
        B, N = points.shape[0], points.shape[1]
        pts = points.detach().to(self.device, torch.float32).contiguous()
        cc = crop_center.to(self.device, torch.float32).contiguous()
        bc = body_center.to(self.device, torch.float32).contiguous()
        out = torch.empty(B, N_OUT, N, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call("vt_query_fwd_tc_heads", _lib.ptr(pts), _lib.ptr(cc), _lib.ptr(bc), B, N, _lib.ptr(im_feat), _lib.ptr(tmpx),
                      _lib.ptr(tri_tmpx), _lib.ptr(tri_feat), im_feat.shape[1], im_feat.shape[2], tmpx.shape[1], tmpx.shape[2],
                      self._cam7, _lib.ptr(self._wpack), *(_lib.ptr(t) for t in self._wtc), mask, _lib.ptr(out), None,
                      _lib.ptr(self._q_overflow), _lib.stream_ptr())
        res = {}
        for h in heads:
            lo, hi = HEAD_SLICES[names.index(h)]
            res[h] = out[:, lo:hi].reshape(B, 3, 3, N) if h == "pca" else out[:, lo:hi]
        return res

    def query(self, points, crop_center=None, **kwargs):
        """CHORETriplane.query (model/chore_triplane.py:97-164).  Stores ``self.preds`` = (df [B,2,N], pca [B,3,3,N],
        parts [B,14,N], centers [B,3,N], visibility [B,1,N]); differentiable w.r.t. ``points``."""
        body_center = kwargs.get("body_center")
        if crop_center is None or body_center is None:
            raise ValueError("query() needs crop_center and body_center (model/chore_triplane.py:114,130)")
        self.points, self.crop_center = points, crop_center
        B, N = points.shape[0], points.shape[1]
        if points.requires_grad and torch.is_grad_enabled():
            df, pca, parts, centers, vis, xy = _QueryFn.apply(self, points, crop_center, body_center)
        else:
            out, xy = self._query_raw(points, crop_center, body_center, want_xy=True)
            df, pca, parts, centers, vis = (out[:, lo:hi] for lo, hi in HEAD_SLICES)
        pca = pca.reshape(B, 3, 3, N)
        self.points_xy = xy
        self.preds = (df, pca, parts, centers, vis)
        self.intermediate_preds_list = [self.preds]
        self.local_feat_list = None

    def query_features(self, points, crop_center=None, **kwargs):
        """CHORETriplane.query_features (model/chore_triplane.py:166-205): ([B, 611, N] features, [B, 2, N] xy)."""
        return self._query_raw(points, crop_center, kwargs.get("body_center"), want_xy=True, want_feat=True)

    def get_preds(self):
        return self.preds

    def project_points(self, points, offsets):
        """KinectColorCamera.project_points (model/camera.py:45-50): [B, 3, N] = (nx, ny, z)."""
        d = self.dims
        x, y, z = points[..., 0], points[..., 1], points[..., 2]
        px = d.crop_size / 2 + (d.fx_px * x / z + d.cx_px) - offsets[:, 0:1]
        py = d.crop_size / 2 + (d.fy_px * y / z + d.cy_px) - offsets[:, 1:2]
        return torch.stack([2 * px / d.crop_size - 1, 2 * py / d.crop_size - 1, z], 1)

    @staticmethod
    def triplane_project(points, body_center, fx=1.0, cx=0.0):
        """model/chore_triplane.py:220-251."""
        c = points - body_center[:, None, :]
        return [torch.stack([c[..., 2] * fx + cx, c[..., 1] * fx + cx], 1),
                torch.stack([-c[..., 0] * fx + cx, c[..., 1] * fx + cx], 1),
                torch.stack([c[..., 0] * fx + cx, -c[..., 2] * fx + cx], 1)]
