"""SO(3) projection and ragged Chamfer on the B200 against their CPU restatements (oracle/geom_ref.py)."""
import pytest
import torch

from conftest import rel_err
from oracle import geom_ref as R

pytestmark = pytest.mark.gpu


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _rand_rot(n, gen):
    q = torch.randn(n, 4, generator=gen, dtype=torch.float64)
    q = q / q.norm(dim=1, keepdim=True)
    w, x, y, z = q.unbind(1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y), 2 * (x * y + w * z), 1 - 2 * (x * x + z * z),
                        2 * (y * z - w * x), 2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], 1).reshape(n, 3, 3)


def test_project_so3_forward_backward():
    _need_gpu()
    from vistracker_b200.geom import project_so3
    gen = torch.Generator().manual_seed(0)
    near_rot = _rand_rot(96, gen) + 1e-2 * torch.randn(96, 3, 3, generator=gen, dtype=torch.float64)   # the fitter's case
    generic = torch.randn(64, 3, 3, generator=gen, dtype=torch.float64)                                 # includes det < 0
    for mats in (near_rot, generic):
        ref_in = mats.clone().requires_grad_(True)
        ref = R.project_so3(ref_in)
        x = mats.float().cuda().requires_grad_(True)
        out = project_so3(x)
        assert rel_err(out.detach().cpu(), ref.detach()) < 1e-5
        eye = torch.eye(3).expand_as(out.detach().cpu())
        assert rel_err(out.detach().cpu() @ out.detach().cpu().transpose(1, 2), eye) < 1e-5
        g = torch.randn(mats.shape, generator=gen, dtype=torch.float64)
        (ref * g).sum().backward(); (out * g.float().cuda()).sum().backward()
        keep = torch.linalg.svdvals(mats)[:, 1:].sum(1) > 0.2          # the Jacobian blows up when s2 + s3 -> 0
        assert rel_err(x.grad.cpu()[keep], ref_in.grad[keep]) < 1e-4


def test_decopose_axis_replays_injected_noise():
    _need_gpu()
    from vistracker_b200.geom import decopose_axis
    gen = torch.Generator().manual_seed(3)
    rot, noise = _rand_rot(8, gen).float(), torch.rand(8, 3, 3, generator=gen)
    ref = R.project_so3((rot + 1e-4 * noise).double())
    assert rel_err(decopose_axis(rot.cuda(), noise=noise.cuda()).cpu(), ref) < 1e-5
    assert rel_err(decopose_axis(rot.cuda(), no_rand=True).cpu(), R.project_so3(rot.double())) < 1e-5


def test_ragged_chamfer_forward_backward():
    _need_gpu()
    from vistracker_b200.geom import chamfer_distance_ragged
    gen = torch.Generator().manual_seed(1)
    sizes = [(1, 1), (5, 300), (257, 3), (64, 64), (700, 150)]
    xs = [torch.randn(a, 3, generator=gen) for a, _ in sizes]
    ys = [torch.randn(b, 3, generator=gen) + 0.3 for _, b in sizes]
    rx = [t.double().requires_grad_(True) for t in xs]; ry = [t.double().requires_grad_(True) for t in ys]
    ref = R.chamfer_ragged(rx, ry); ref.backward()
    gx = [t.cuda().requires_grad_(True) for t in xs]; gy = [t.cuda().requires_grad_(True) for t in ys]
    out = chamfer_distance_ragged(gx, gy); out.backward()
    assert abs(float(out) - float(ref)) < 1e-5 * abs(float(ref))
    for a, b in zip(gx + gy, rx + ry):
        assert rel_err(a.grad.cpu(), b.grad) < 1e-5


def test_eval_chamfer_matches_reference_golden_and_oracle():
    """vt_nn_dist through geom.eval_chamfer_distance: the reference function's values (sklearn kd-tree, eval_chamfer.npz), all three
    directions, single clouds and a batch of BASELINE-sized clouds (10 000 samples) against the float64 restatement."""
    import os
    import numpy as np
    from oracle import geom_ref as GR
    from vistracker_b200.geom import eval_chamfer_distance
    _need_gpu()
    dev = lambda: torch.device("cuda", 0)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "eval_chamfer.npz"))
    for i in range(3):
        x, y = torch.from_numpy(gold[f"x{i}"]).to(dev()), torch.from_numpy(gold[f"y{i}"]).to(dev())
        for d in ("bi", "x_to_y", "y_to_x"):
            got, ref = float(eval_chamfer_distance(x, y, d)), float(gold[f"cd{i}_{d}"])
            assert abs(got - ref) <= 2e-6 * max(1.0, abs(ref)), (i, d, got, ref)
    g = torch.Generator().manual_seed(5)
    xb = torch.randn(3, 10000, 3, generator=g) * 0.5
    yb = xb[:, torch.randperm(10000, generator=g)[:9000]] + 0.01 * torch.randn(3, 9000, 3, generator=g)
    got = eval_chamfer_distance(xb.to(dev()), yb.to(dev())).cpu()
    for b in range(3):
        ref = GR.eval_chamfer(xb[b, ::1].numpy()[:2000], yb[b].numpy(), "x_to_y")          # float64 brute force on a slice (memory)
        sub = float(eval_chamfer_distance(xb[b, :2000].to(dev()), yb[b].to(dev()), "x_to_y"))
        assert abs(sub - ref) <= 2e-6 * max(1.0, ref)
    assert got.shape == (3,) and bool((got > 0).all())
    with pytest.raises(ValueError, match="Invalid direction"):
        eval_chamfer_distance(xb[0].to(dev()), yb[0].to(dev()), "xy")


def test_procrustes_matches_reference_golden():
    """vt_procrustes / vt_similarity_apply against compute_transform / compute_similarity_transform of recon/eval/pose_utils.py (goldens),
    including a reflected target (determinant fix) and a batch."""
    import os
    import numpy as np
    from vistracker_b200.geom import apply_similarity, procrustes_transform
    _need_gpu()
    dev = torch.device("cuda", 0)
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "eval_chamfer.npz"))
    for i in range(2):
        src, dst = torch.from_numpy(gold[f"pa_src{i}"]).to(dev), torch.from_numpy(gold[f"pa_dst{i}"]).to(dev)
        R, t, s = procrustes_transform(src, dst)
        assert np.abs(R.cpu().numpy() - gold[f"pa_R{i}"]).max() < 2e-5
        assert np.abs(t.cpu().numpy() - gold[f"pa_t{i}"]).max() < 5e-5 and abs(float(s) - float(gold[f"pa_s{i}"])) < 2e-5
        assert abs(float(torch.linalg.det(R.double().cpu())) - 1.0) < 1e-5
        hat = apply_similarity(src, R, t, s)
        assert rel_err(hat.cpu(), gold[f"pa_hat{i}"]) < 2e-5
    a = torch.stack([torch.from_numpy(gold["pa_src0"]), torch.from_numpy(gold["pa_src0"]).flip(0)]).to(dev)
    b = torch.stack([torch.from_numpy(gold["pa_dst0"]), torch.from_numpy(gold["pa_dst0"]).flip(0)]).to(dev)
    Rb, tb, sb = procrustes_transform(a, b)
    assert rel_err(Rb[0].cpu(), Rb[1].cpu()) < 1e-5 and rel_err(Rb[0].cpu(), gold["pa_R0"]) < 2e-5


def test_pca_orientation_matches_reference_golden_and_oracle():
    """vt_pca_orientation: PCAUtil.init_object_orientation's values (infill_small.npz), a per-frame template stack, and the decopose_axis
    variant with injected noise against the float64 restatement."""
    import os
    import numpy as np
    from oracle import geom_ref as GR
    from vistracker_b200.geom import init_object_orientation
    _need_gpu()
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "infill_small.npz"))
    tgt, src = torch.from_numpy(g["pca_tgt"]).cuda(), torch.from_numpy(g["pca_src"]).cuda()
    got = init_object_orientation(tgt, src, no_rand=True)
    assert np.abs(got.cpu().numpy() - g["pca_R"]).max() < 2e-5
    stack = src[None].repeat(tgt.shape[0], 1, 1)
    assert torch.equal(init_object_orientation(tgt, stack, no_rand=True), got)
    noise = torch.rand(tgt.shape[0], 3, 3, generator=torch.Generator().manual_seed(2))
    ref = GR.init_object_orientation(tgt.cpu(), src.cpu(), noise=noise)
    assert rel_err(init_object_orientation(tgt, src, noise=noise.cuda()).cpu(), ref) < 1e-5
    assert init_object_orientation(tgt, src).shape == (tgt.shape[0], 3, 3)
    with pytest.raises(ValueError, match="invalid shapes"):
        init_object_orientation(tgt, stack[:3])
