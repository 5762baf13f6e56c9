"""CPU restatement of the UDF -> point-cloud generator (TEST INFRASTRUCTURE).

Follows recon/gen/generator.py:72-104 (approx_surface), :149-215 (gen_pc_batch), generator_triplane.py:32-55 and
generator_vis.py:19-56 line by line, over the oracle's SIF-Net (oracle/sifnet_ref.py) instead of the reference nn.Module; the
random draws come from torch's CPU generator in the reference's order.  Pinned by tests/golden/generator_small.npz: the reference's
own GeneratorTriplaneVis.get_grid_samples / approx_surface / gen_pc_batch executed on the CPU (the instance is created without
``__init__``, which only wants a checkpoint directory and a CUDA device, generator.py:28-52) -- this restatement reproduces them
bit for bit, including the sample counts and the resampling draws (tests/test_oracle_generator.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import sifnet_ref as R


def approx_surface(sd, maps, samples, num_steps, crop, body, cam, df_idx, threshold):
    preds = None
    for _ in range(num_steps):
        samples = samples.detach().requires_grad_(True)
        preds = R.sif_query(sd, maps, samples, crop, body, cam)
        df_target = torch.clamp(preds[0][:, df_idx, :], max=threshold)
        df_target.sum().backward()
        gradient = samples.grad.detach()
        samples = (samples.detach() - F.normalize(gradient, dim=2) * df_target.detach().unsqueeze(-1)).detach()
    return samples, [p.detach() for p in preds]


def gen_pc_batch(sd, maps, df_type, samples_init, num_points, crop, body, cam, num_steps, threshold=2.0, filter_val=0.004,
                 sample_num=20000, max_iter=100):
    df_idx = 0 if df_type == "human" else 1
    B = samples_init.shape[0]
    names = ["points", "pca_axis", "parts", "centers", "visibility"]
    out = {n: [[] for _ in range(B)] for n in names}
    it, count = 0, 0
    samples = samples_init.clone()
    while count < num_points:
        surf, preds = approx_surface(sd, maps, samples, num_steps, crop, body, cam, df_idx, threshold)
        df_target = torch.clamp(preds[0][:, df_idx, :], max=threshold)
        mask = (df_target < filter_val) & (surf[:, :, 2] > 1.0)
        if it > 0:
            counts = []
            for i in range(B):
                out["points"][i].append(surf[i, mask[i]])
                for n, p in zip(names[1:], preds[1:]):
                    out[n][i].append(p[i, ..., mask[i]])
                counts.append(int(mask[i].sum()))
            count += int(np.min(counts))
        new = []
        for i in range(B):
            s_i = samples[i, mask[i], :].unsqueeze(0)
            if s_i.shape[1] > 1:
                idx = torch.randint(s_i.shape[1], (1, sample_num))
                s_i = s_i[[[0, ] * sample_num], idx]
                s_i = s_i + (threshold / 3) * torch.randn(s_i.shape)
            else:
                idx = torch.randint(samples_init.shape[1], (1, sample_num))
                s_i = samples_init[[[i, ] * sample_num], idx].clone()
                s_i = s_i + 0.5 * torch.randn(1, sample_num, 3)
            new.append(s_i)
        samples = torch.cat(new, 0)
        it += 1
        if it == max_iter:
            raise RuntimeError("generation failed")
    res = {}
    for n in names:
        comb = []
        for i in range(B):
            if n == "points":
                comb.append(torch.cat(out[n][i], 0)[:count]); continue
            o = torch.cat(out[n][i], -1)[..., :count]
            comb.append(torch.argmax(o, 0) if n == "parts" else torch.mean(o, -1))
        res[n] = torch.stack(comb, 0)
    res["centers"] = torch.cat([torch.full_like(res["centers"], float("nan")), res["centers"]], 1)
    return res
