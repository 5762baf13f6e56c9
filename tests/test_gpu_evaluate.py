"""Sequence evaluation (window Procrustes + Chamfer + v2v) on the B200 against the CPU restatement of the reference's loop."""
import numpy as np
import pytest
import torch

from oracle import geom_ref as G

pytestmark = pytest.mark.gpu


def _sequence(T=23, Vs=300, Vo=120, seed=0):
    rng = np.random.default_rng(seed)
    base_s, base_o = rng.standard_normal((Vs, 3)) * [0.25, 0.8, 0.2], rng.standard_normal((Vo, 3)) * 0.3 + [0.6, 0.0, 0.1]
    t = np.arange(T)[:, None, None] / 10.0
    gt_s = base_s[None] + np.concatenate([0.3 * np.sin(t), 0 * t, 2.5 + 0.1 * t], -1)
    gt_o = base_o[None] + np.concatenate([0.3 * np.sin(t), 0.05 * t, 2.5 + 0.1 * t], -1)
    ang = 0.3
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    rec_s = 1.1 * gt_s.dot(R.T) + [0.2, -0.1, 0.4] + 0.01 * rng.standard_normal(gt_s.shape)
    rec_o = 1.1 * gt_o.dot(R.T) + [0.2, -0.1, 0.4] + 0.02 * rng.standard_normal(gt_o.shape)
    return [a.astype(np.float32) for a in (rec_s, rec_o, gt_s, gt_o)]


@pytest.mark.parametrize("window,smpl_only", [(7, False), (300, True), (0, False)])
def test_evaluate_sequence_matches_restatement(window, smpl_only):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from vistracker_b200.evaluate import evaluate_sequence
    rec_s, rec_o, gt_s, gt_o = _sequence()
    exist = np.ones(len(gt_s), bool)
    exist[[3, 13, 14]] = False
    ref = G.evaluate_sequence(rec_s, rec_o, gt_s, gt_o, window, exist, smpl_only)
    dev = torch.device("cuda", 0)
    errs, keep, transforms = evaluate_sequence(*(torch.from_numpy(a).to(dev) for a in (rec_s, rec_o, gt_s, gt_o)), window=window,
                                               recon_exist=exist, sample_num=None, smpl_only=smpl_only)
    assert keep == [i for i in range(len(gt_s)) if exist[i]]
    assert errs.shape == ref.shape
    assert np.abs(errs.cpu().numpy() - ref).max() < 2e-3 * max(1.0, np.abs(ref).max())          # cm; fp32 kernels vs float64 numpy
    if window == 7:
        assert [t[0] for t in transforms] == [0, 6, 13, 20]          # the reference's schedule: frame 0, then every count % window == 0


def test_surface_samples_lie_on_the_mesh():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from vistracker_b200.evaluate import sample_surface
    dev = torch.device("cuda", 0)
    verts = torch.tensor([[[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]]], device=dev)
    faces = torch.tensor([[0, 1, 2], [1, 3, 2]], device=dev)
    pts = sample_surface(verts, faces, 4096, torch.Generator(device=dev).manual_seed(0))
    assert pts.shape == (1, 4096, 3) and float(pts[..., 2].abs().max()) == 0.0
    assert float(pts[..., :2].min()) >= 0.0 and float(pts[..., :2].max()) <= 1.0
    assert abs(float((pts[..., 0] + pts[..., 1] < 1).float().mean()) - 0.5) < 0.05            # both triangles get half of the samples
