"""Neural-UDF -> surface point cloud (``Generator`` / ``GeneratorTriplane`` / ``GeneratorTriplaneVis``,
recon/gen/generator.py:72-257, generator_triplane.py:32-55, generator_vis.py:15-56) over the fused B200 query kernels.

One projection step of ``approx_surface`` (query, ``df.sum().backward()``, ``p -= normalize(grad) * df``) is ONE kernel launch
(``vt_query_project_step``): the 9 non-final steps evaluate only the distance head, the final one all five heads.
The reference draws every random number from torch's CPU generator (``torch.rand / randint / randn`` without a device) and
moves it to the GPU; this module draws the same tensors in the same order, so with equal seeds both implementations see
identical samples.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

from . import _lib
from .sifnet import CHORETriplaneVisibility, N_OUT

P, S = _lib.ptr, _lib.stream_ptr


class GeneratorTriplaneVis:
    def __init__(self, model: CHORETriplaneVisibility, threshold=2.0, sparse_thres=0.03, filter_val=0.004, device=None):
        self.model = model
        self.device = torch.device(device) if device is not None else model.device
        self.threshold, self.sparse_thres, self.filter_val = threshold, sparse_thres, filter_val
        self.sample_num = 100000

    # ------------------------------------------------------------------ pieces with the reference's names
    def filter(self, data):
        self.model.filter(data["images"].to(self.device))

    def prep_query_input(self, batch):
        return {"crop_center": batch["crop_center"].to(self.device), "body_center": batch["body_center"].to(self.device)}

    def get_grid_samples(self, sample_num, batch_size=1, body_center=None):
        """generator_triplane.py:32-55: U([-1,1] x [-1.5,1.5] x [-.6,.6]) around the body centre (CPU generator draw)."""
        assert body_center is not None
        samples = torch.rand(batch_size, sample_num, 3).float().to(self.device)
        samples[:, :, 0] = samples[:, :, 0] * 2 - 1
        samples[:, :, 1] = samples[:, :, 1] * 3 - 1.5
        samples[:, :, 2] = samples[:, :, 2] * 1.2 - 0.6
        return samples + body_center.unsqueeze(1).to(self.device)

    @staticmethod
    def get_out_names():
        return ["points", "pca_axis", "parts", "centers", "visibility"]

    def _project_step(self, pts, query_input, df_idx, want_preds):
        net = self.model
        im_feat, tmpx, tri_tmpx, tri_feat = net._maps
        B, N = pts.shape[0], pts.shape[1]
        new = torch.empty_like(pts)
        out = torch.empty(B, N_OUT, N, device=self.device) if want_preds else None
        d = net.dims
        with torch.cuda.device(self.device):
            if not net.query_on_cuda_cores:
                # tensor-core path (csrc/query_bwd_tc.cu): the step evaluates the distance head only; the predictions of the final
                # step come from one forward launch at the same points
                _lib.call("vt_query_project_step_tc", P(pts), P(query_input["crop_center"]), P(query_input["body_center"]), B, N,
                          P(im_feat), P(tmpx), P(tri_tmpx), P(tri_feat), im_feat.shape[1], im_feat.shape[2], tmpx.shape[1],
                          tmpx.shape[2], net._cam7, P(net._wpack), *(P(t) for t in net._wtc), *(P(t) for t in net._wtc_bwd), df_idx,
                          float(self.threshold), P(new), None, P(net._q_overflow), S())
                if want_preds:
                    _lib.call("vt_query_fwd_tc", P(pts), P(query_input["crop_center"]), P(query_input["body_center"]), B, N, P(im_feat),
                              P(tmpx), P(tri_tmpx), P(tri_feat), im_feat.shape[1], im_feat.shape[2], tmpx.shape[1], tmpx.shape[2],
                              net._cam7, P(net._wpack), *(P(t) for t in net._wtc), P(out), None, P(net._q_overflow), S())
                return new, out
            _lib.call("vt_query_project_step", P(pts), P(query_input["crop_center"]), P(query_input["body_center"]), B, N, P(im_feat),
                      P(tmpx), P(tri_tmpx), P(tri_feat), im_feat.shape[1], im_feat.shape[2], tmpx.shape[1], tmpx.shape[2], d.rgb.out_ch,
                      d.rgb.stem_ch, d.tri.stem_ch, d.tri.out_ch, net._cam7, P(net._wpack), P(net._wpack_bwd), df_idx,
                      float(self.threshold), P(new), P(out), None, S())
        return new, out

    def approx_surface(self, samples, num_steps, query_input, df_type):
        """generator.py:72-104.  Returns (projected samples, predictions of the LAST query, i.e. at the positions before the
        final update) -- exactly what the reference returns."""
        df_idx = 0 if df_type == "human" else 1
        pts = samples.detach().float().contiguous()
        qi = {k: v.float().contiguous() for k, v in query_input.items()}
        out = None
        for j in range(num_steps):
            pts, out = self._project_step(pts, qi, df_idx, want_preds=(j == num_steps - 1))
        B, N = pts.shape[0], pts.shape[1]
        preds = (out[:, 0:2], out[:, 2:11].reshape(B, 3, 3, N), out[:, 11:25], out[:, 25:28], out[:, 28:29])
        return pts, preds

    def gen_pc_batch(self, df_type, samples_init, num_points, batch, num_steps, max_iter=100, mute=True):
        """generator.py:149-215."""
        query_input = self.prep_query_input(batch)
        df_idx = 0 if df_type == "human" else 1
        batch_size = samples_init.shape[0]
        out_names = self.get_out_names()
        out_dict: Dict[str, List[list]] = {n: [[] for _ in range(batch_size)] for n in out_names}
        sample_num = 20000
        it, samples_count = 0, 0
        samples = samples_init.clone().to(self.device)
        n_init = samples_init.shape[1]
        samples_init_dev = samples_init.to(self.device)
        while samples_count < num_points:
            samples_surface, preds = self.approx_surface(samples, num_steps, query_input, df_type)
            df_target = torch.clamp(preds[0][:, df_idx, :], max=self.threshold)
            mask = (df_target < self.filter_val) & (samples_surface[:, :, 2] > 1.0)
            counts_dev = mask.sum(1)
            order = torch.argsort((~mask).to(torch.uint8), dim=1, stable=True)
            # While the device works on the launches above: the noise of the resampling below, drawn from torch's CPU generator in the
            # reference's order (per frame: randint(high, (1, sample_num)), then randn(1, sample_num, 3)).  The noise does not depend on the
            # survivor counts, the randint range does -- but randint consumes the same generator outputs whatever its range (one 32-bit draw
            # per element below 2^32), so a dummy randint advances the generator here and the real one is drawn after the counts are known,
            # from the state saved in front of it.  Identical samples for equal seeds, and the 0.22 ms per frame of randn no longer sit
            # between the round's host sync and the next round's first launch.
            states, noise = [], torch.empty(batch_size, sample_num, 3, pin_memory=True)
            for i in range(batch_size):
                states.append(torch.get_rng_state())
                torch.randint(2, (1, sample_num))
                noise[i] = torch.randn(1, sample_num, 3)[0]
            final_state = torch.get_rng_state()
            # ONE host sync per round: the per-frame counts.  `order` lists the surviving sample positions of every frame first (stable),
            # so frame i's survivors are order[i, :counts[i]] -- the boolean-mask indexing of the reference without a sync per frame
            counts = counts_dev.cpu().tolist()
            if it > 0:
                for i in range(batch_size):
                    keep = order[i, :counts[i]]
                    out_dict["points"][i].append(samples_surface[i].index_select(0, keep))
                    for name, pred in zip(out_names[1:], preds[1:]):
                        out_dict[name][i].append(pred[i].index_select(-1, keep))
                samples_count += int(np.min(counts))
                if not mute:
                    print("{} points".format(samples_count))
            indices = torch.empty(batch_size, sample_num, dtype=torch.int64, pin_memory=True)
            for i in range(batch_size):
                torch.set_rng_state(states[i])
                indices[i] = torch.randint(counts[i] if counts[i] > 1 else n_init, (1, sample_num))[0]
            torch.set_rng_state(final_state)
            indices_dev, noise_dev = indices.to(self.device, non_blocking=True), noise.to(self.device, non_blocking=True)
            samples_new = []
            for i in range(batch_size):
                if counts[i] > 1:       # around the survivors: sigma = threshold / 3 (the product is formed in fp32 as on the host)
                    src = samples[i].index_select(0, order[i, :counts[i]].index_select(0, indices_dev[i]))
                    samples_i = src.unsqueeze(0) + ((self.threshold / 3) * noise_dev[i]).unsqueeze(0)
                else:                   # no survivor: around the initial grid samples, sigma = 0.5
                    src = samples_init_dev[i].index_select(0, indices_dev[i])
                    samples_i = src.unsqueeze(0) + (0.5 * noise_dev[i]).unsqueeze(0)
                samples_new.append(samples_i)
            samples = torch.cat(samples_new, 0).detach()
            it += 1
            if it == max_iter:
                raise RuntimeError(f"point generation for df {df_type} failed after {max_iter} iterations")
        return self.compose_outdict(batch_size, out_dict, out_names, samples_count)

    @staticmethod
    def compose_outdict(batch_size, out_dict, out_names, samples_count):
        """generator_vis.py:19-56: truncate to the common count, argmax parts, average pca / centres / visibility, prepend NaNs."""
        for name in out_names:
            comb = []
            for i in range(batch_size):
                if name == "points":
                    comb.append(torch.cat(out_dict[name][i], 0)[:samples_count, :])
                    continue
                o = torch.cat(out_dict[name][i], -1)[..., :samples_count]
                if name == "parts":
                    o = torch.argmax(o, 0)
                elif name == "pca_axis":
                    o = torch.mean(o, -1)
                elif name in ("centers", "visibility"):
                    o = torch.mean(o, -1)
                comb.append(o)
            out_dict[name] = torch.stack(comb, 0).cpu()           # the reference collects on the host (generator.py:119-125)
        nan_values = torch.zeros_like(out_dict["centers"]) + float("nan")
        out_dict["centers"] = torch.cat([nan_values, out_dict["centers"]], 1)
        return out_dict

    def generate_pclouds_batch(self, data, num_steps=10, num_points=50000, mute=True):
        """generator.py:127-147."""
        self.filter(data)
        batch_size = data["images"].shape[0]
        samples = self.get_grid_samples(30000, batch_size=batch_size, body_center=data.get("body_center"))
        return {t: self.gen_pc_batch(t, samples, num_points, data, num_steps, mute=mute) for t in ("human", "object")}
