"""oracle/geom_ref.py: eval_chamfer against the reference's own recon/eval/chamfer_distance.py (eval_chamfer.npz)."""
import os

import numpy as np

from oracle import geom_ref as G

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "eval_chamfer.npz"))


def test_eval_chamfer_matches_reference():
    for i in range(3):
        for d in ("bi", "x_to_y", "y_to_x"):
            ref = float(GOLD[f"cd{i}_{d}"])
            assert abs(G.eval_chamfer(GOLD[f"x{i}"], GOLD[f"y{i}"], d) - ref) <= 1e-9 * max(1.0, abs(ref)), (i, d)


def test_compute_transform_matches_reference():
    for i in range(2):
        R, t, s = G.compute_transform(GOLD[f"pa_src{i}"], GOLD[f"pa_dst{i}"])
        assert np.abs(R - GOLD[f"pa_R{i}"]).max() < 1e-10 and np.abs(t - GOLD[f"pa_t{i}"]).max() < 1e-10 and abs(s - float(GOLD[f"pa_s{i}"])) < 1e-10
        hat = s * GOLD[f"pa_src{i}"].astype(np.float64).dot(R.T) + t
        assert np.abs(hat - GOLD[f"pa_hat{i}"]).max() < 1e-9
