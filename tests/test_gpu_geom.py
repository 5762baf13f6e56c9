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
