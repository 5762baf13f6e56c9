"""oracle/generator_ref.py against the reference's own generator methods run on the CPU (tests/golden/generator_small.npz: grid samples,
three chained projection steps for both targets, and a whole gen_pc_batch with its resampling draws) -- CPU only."""
import os

import numpy as np
import torch

from oracle import generator_ref as G
from oracle import sifnet_ref as R
from vistracker_b200 import default_options, resolve_dims
from vistracker_b200.synth import synthetic_frames, synthetic_state_dict

GOLD = os.path.join(os.path.dirname(__file__), "golden", "generator_small.npz")
DIMS = resolve_dims(default_options())
CAM = (DIMS.fx_px, DIMS.fy_px, DIMS.cx_px, DIMS.cy_px, DIMS.crop_size)


def _close(a, b, frac_tol=1e-4, worst=2e-2):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    err = (a - b).abs().reshape(a.shape[0], -1).max(-1).values / b.abs().max()
    return float((err < frac_tol).double().mean()), float(err.max())


def test_generator_restatement_matches_reference_methods():
    g = np.load(GOLD)
    sd = synthetic_state_dict(DIMS, seed=0)
    images, _, crop, body = synthetic_frames(2, size=64, seed=11, n_points=4, jitter=True)
    with torch.no_grad():
        maps = R.sif_filter(sd, images)
    torch.manual_seed(123)
    init = torch.rand(2, 300, 3).float()                                         # generator_triplane.py:46-53, same draw order
    init = init * torch.tensor([2.0, 3.0, 1.2]) - torch.tensor([1.0, 1.5, 0.6]) + body.unsqueeze(1)
    assert np.abs(init.numpy() - g["init"]).max() < 1e-6
    for name, idx in (("human", 0), ("object", 1)):
        surf, preds = G.approx_surface(sd, maps, torch.from_numpy(g["init"]), 3, crop, body, CAM, idx, 2.0)
        frac, worst = _close(surf.reshape(-1, 1, 3), torch.from_numpy(g[f"surf_{name}"]).reshape(-1, 1, 3))
        assert frac > 0.97 and worst < 2e-2, (name, frac, worst)                  # bit-identical on the machine that wrote the golden; another CPU's rounding is amplified by the nearly flat random UDF
        assert np.abs(preds[0].numpy() - g[f"surf_df_{name}"]).max() < 2e-3 * np.abs(g[f"surf_df_{name}"]).max()
    torch.manual_seed(7)
    pc = G.gen_pc_batch(sd, maps, "object", torch.from_numpy(g["init"]), 250, crop, body, CAM, num_steps=1, filter_val=10.0)
    assert tuple(pc["points"].shape) == g["pc_points"].shape                      # same masks, same counts, same resampling draws
    frac, worst = _close(pc["points"].reshape(-1, 1, 3), torch.from_numpy(g["pc_points"]).reshape(-1, 1, 3))
    assert frac > 0.97 and worst < 5e-2, (frac, worst)
    assert np.abs(pc["pca_axis"].numpy() - g["pc_pca_axis"]).max() < 1e-3 and np.abs(pc["visibility"].numpy() - g["pc_visibility"]).max() < 1e-3
    assert np.isnan(g["pc_centers"][:, :3]).all() and torch.isnan(pc["centers"][:, :3]).all()
    assert np.abs(pc["centers"][:, 3:].numpy() - g["pc_centers"][:, 3:]).max() < 1e-3
    assert float((pc["parts"].numpy() == g["pc_parts"]).mean()) > 0.99
