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
