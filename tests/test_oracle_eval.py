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


def test_evaluation_loop_matches_reference_eva_seq():
    """oracle evaluate_sequence (alignment windows, missing reconstructions, Chamfer on vertices, v2v, acceleration errors) against the
    reference's VideoPackedEvaluator.eva_seq run on in-memory arrays (tests/golden/eval_seq.npz)."""
    import os
    import numpy as np
    from oracle import geom_ref as GR
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "eval_seq.npz"))
    W = int(g["window"])
    for tag, exist in (("all", None), ("gaps", g["exist"])):
        got = GR.evaluate_sequence(g["sv_rc"], g["ov_rc"], g["sv_gt"], g["ov_gt"], W, recon_exist=exist, with_accel=True)
        ref = g[f"errors_{tag}"]
        assert got.shape == ref.shape, tag
        assert np.allclose(got, ref, rtol=1e-7, atol=1e-8, equal_nan=True), (tag, np.nanmax(np.abs(got - ref)))
    assert g["errors_gaps"].shape[0] == int(g["exist"].sum()) and g["errors_all"].shape == (23, 6)
    four = GR.evaluate_sequence(g["sv_rc"], g["ov_rc"], g["sv_gt"], g["ov_gt"], W, recon_exist=g["exist"])
    assert np.allclose(four, g["errors_gaps"][:, :4], rtol=1e-7, atol=1e-8)


def test_acceleration_error_helper_matches_reference_columns():
    """vistracker_b200.evaluate.acceleration_errors (plain tensor arithmetic, runs on any device) with the oracle's window alignments."""
    import os
    import numpy as np
    import torch
    from oracle import geom_ref as GR
    from vistracker_b200.evaluate import acceleration_errors
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "eval_seq.npz"))
    W = int(g["window"])
    for tag, exist in (("all", None), ("gaps", g["exist"])):
        _, tr = GR.evaluate_sequence(g["sv_rc"], g["ov_rc"], g["sv_gt"], g["ov_gt"], W, recon_exist=exist, return_transforms=True)
        acc_s = acceleration_errors(torch.from_numpy(g["sv_rc"]), torch.from_numpy(g["sv_gt"]), tr, W, exist)
        acc_o = acceleration_errors(torch.from_numpy(g["ov_rc"]), torch.from_numpy(g["ov_gt"]), tr, W, exist)
        ref = g[f"errors_{tag}"]
        assert np.allclose(acc_s.numpy(), ref[:, 4], rtol=1e-7, atol=1e-9, equal_nan=True), tag
        assert np.allclose(acc_o.numpy(), ref[:, 5], rtol=1e-7, atol=1e-9, equal_nan=True), tag


def test_chamfer_ragged_agrees_with_an_independent_kd_tree_computation():
    """pytorch3d.loss.chamfer_distance(Pointclouds(x), Pointclouds(y)) -- the operator behind compute_contact_loss
    (recon/recon_fit_trivis_full.py:393-457) -- is un-vendored, so oracle/geom_ref.chamfer_ragged restates its published
    definition: per cloud pair the mean SQUARED nearest-neighbour distance in both directions (point_reduction='mean'), averaged over the
    batch (batch_reduction='mean').  Cross-check against a computation that shares no code with it: exact nearest neighbours from
    scikit-learn's KDTree (the structure the reference's own evaluation uses, recon/eval/chamfer_distance.py:28-33) in float64, on ragged
    clouds including a single-point cloud and duplicated points."""
    import torch
    from sklearn.neighbors import KDTree
    rng = np.random.default_rng(5)
    sizes = [(37, 210), (1, 64), (150, 3), (96, 96)]
    xs = [rng.standard_normal((a, 3)) * 0.3 for a, _ in sizes]
    ys = [rng.standard_normal((b, 3)) * 0.3 + 0.1 for _, b in sizes]
    xs[3][10:20] = xs[3][0]                                                        # duplicates: ties between neighbours
    ref = 0.0
    for x, y in zip(xs, ys):
        dxy = KDTree(y).query(x, k=1)[0][:, 0]
        dyx = KDTree(x).query(y, k=1)[0][:, 0]
        ref += (dxy ** 2).mean() + (dyx ** 2).mean()
    ref /= len(xs)
    got = G.chamfer_ragged([torch.from_numpy(x) for x in xs], [torch.from_numpy(y) for y in ys])
    assert abs(float(got) - ref) <= 1e-12 * max(1.0, abs(ref))
    got32 = G.chamfer_ragged([torch.from_numpy(x).float() for x in xs], [torch.from_numpy(y).float() for y in ys])
    assert abs(float(got32) - ref) <= 2e-6 * abs(ref)
