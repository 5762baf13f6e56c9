"""Per-kernel parity on the B200: every C-ABI operator against a plain PyTorch fp32/fp64 CPU reference of the same op."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL = 2e-5      # single ops; the end-to-end bar (north_star) is 1e-4


def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda", 0)


def nhwc(x):    # NCHW cpu -> NHWC cuda
    return x.permute(0, 2, 3, 1).contiguous().to(dev())


def nchw(x):    # NHWC cuda -> NCHW cpu float64
    return x.permute(0, 3, 1, 2).double().cpu()


def chan_stats(x_nchw):
    x = x_nchw.double()
    return torch.stack([x.sum((2, 3)), (x * x).sum((2, 3))], -1)


def test_stem_conv_and_stats():
    from vistracker_b200 import ops
    g = torch.Generator().manual_seed(1)
    img = torch.rand(2, 8, 64, 96, generator=g)
    for cin, cout, c_off, views in ((5, 64, 0, 1), (1, 32, 5, 3)):
        w = torch.randn(cout, cin, 7, 7, generator=g) * 0.05
        b = torch.randn(cout, generator=g) * 0.1
        st = ops.new_stats(2 * views, cout, dev())
        out = ops.stem_conv(img.to(dev()), w, b.to(dev()), c_off, cin, views, st)
        for v in range(views):
            ref = F.conv2d(img[:, c_off + v * cin:c_off + (v + 1) * cin].double(), w.double(), b.double(), stride=2, padding=3)
            assert rel_err(nchw(out[v * 2:(v + 1) * 2]), ref) < TOL
            assert rel_err(st[v * 2:(v + 1) * 2].cpu(), chan_stats(ref)) < 1e-5


@pytest.mark.parametrize("C", [32, 64, 128, 256])
def test_groupnorm_affine_act(C):
    from vistracker_b200 import ops
    g = torch.Generator().manual_seed(C)
    x = torch.randn(3, C, 16, 24, generator=g) * 2 + 0.5
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    st = chan_stats(x).to(dev())
    sc, sh = ops.gn_finalize(st, gamma.to(dev()), beta.to(dev()), 16 * 24)
    st2 = ops.new_stats(3, C, dev())
    out = ops.affine_act(nhwc(x), sc, sh, True, st2)
    ref = F.relu(F.group_norm(x.double(), 32, gamma.double(), beta.double(), 1e-5))
    assert rel_err(nchw(out), ref) < TOL
    assert rel_err(st2.cpu(), chan_stats(ref)) < 1e-5


def test_pool_add_upsample():
    from vistracker_b200 import ops
    g = torch.Generator().manual_seed(7)
    for C in (64, 256):
        x = torch.randn(2, C, 16, 32, generator=g)
        y = torch.randn(2, C, 16, 32, generator=g)
        st = ops.new_stats(2, C, dev())
        out = ops.avgpool2(nhwc(x), st)
        ref = F.avg_pool2d(x.double(), 2, stride=2)
        assert rel_err(nchw(out), ref) < 1e-6 and rel_err(st.cpu(), chan_stats(ref)) < 1e-5
        st = ops.new_stats(2, C, dev())
        out = ops.add(nhwc(x), nhwc(y), st)
        assert rel_err(nchw(out), (x + y).double()) < 1e-6 and rel_err(st.cpu(), chan_stats(x + y)) < 1e-5
        low = torch.randn(2, C, 8, 16, generator=g)
        st = ops.new_stats(2, C, dev())
        out = ops.upsample2x_add(nhwc(low), nhwc(x), st)
        ref = x.double() + F.interpolate(low.double(), scale_factor=2, mode="bicubic", align_corners=True)
        assert rel_err(nchw(out), ref) < 1e-5 and rel_err(st.cpu(), chan_stats(ref)) < 1e-5


@pytest.mark.parametrize("ks,cin,cout,H,W", [(3, 64, 64, 12, 20), (3, 32, 32, 8, 8), (1, 128, 256, 7, 9), (3, 256, 128, 16, 16)])
def test_conv_ffma(ks, cin, cout, H, W):
    from vistracker_b200 import ops
    g = torch.Generator().manual_seed(ks * 1000 + cin + cout)
    x = torch.randn(2, cin, H, W, generator=g)
    w = torch.randn(cout, cin, ks, ks, generator=g) * 0.05
    bias = torch.randn(cout, generator=g)
    res = torch.randn(2, cout, H, W, generator=g)
    gamma, beta = torch.randn(cin, generator=g), torch.randn(cin, generator=g)
    sc, sh = ops.gn_finalize(chan_stats(x).to(dev()), gamma.to(dev()), beta.to(dev()), H * W)
    st = ops.new_stats(2, cout, dev())
    out = ops.conv_ffma(nhwc(x), sc, sh, True, w, bias.to(dev()), nhwc(res), stats=st)
    a = F.relu(F.group_norm(x.double(), 32, gamma.double(), beta.double(), 1e-5))
    ref = F.conv2d(a, w.double(), bias.double(), padding=ks // 2) + res.double()
    assert rel_err(nchw(out), ref) < TOL
    assert rel_err(st.cpu(), chan_stats(ref)) < 1e-5
    # no affine / no relu / no bias / no residual, written into a channel slice of a wider tensor
    wide = torch.zeros(2, H, W, cout + 32, device=dev())
    ops.conv_ffma(nhwc(x), None, None, False, w, out=wide[..., 32:])
    assert rel_err(nchw(wide[..., 32:]), F.conv2d(x.double(), w.double(), padding=ks // 2)) < TOL
    assert float(wide[..., :32].abs().max()) == 0.0


def test_prep_split_planes():
    from vistracker_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 32, 8, 16, generator=g) * 3
    sc = torch.rand(2, 32, generator=g) + 0.5
    sh = torch.randn(2, 32, generator=g)
    planes, ovf = ops.prep_split(nhwc(x), sc.to(dev()), sh.to(dev()), True, 1)
    assert tuple(planes.shape) == (2, 2, 10, 18, 64) and int(ovf.item()) == 0
    y = F.relu(x * sc[:, :, None, None] + sh[:, :, None, None])
    rec = planes[0].float() + planes[1].float() / 2048.0
    assert float(rec[:, 0].abs().max()) == 0 and float(rec[:, -1].abs().max()) == 0          # zero border rows
    assert float(rec[:, :, 0].abs().max()) == 0 and float(rec[:, :, -1].abs().max()) == 0    # zero border columns
    assert float(rec[..., 32:].abs().max()) == 0                                             # zero channel padding
    inner = rec[:, 1:-1, 1:-1, :32].permute(0, 3, 1, 2).cpu()
    assert rel_err(inner, y) < 2e-6
    _, ovf = ops.prep_split(nhwc(x * 1e5), None, None, False, 0)
    assert int(ovf.item()) > 0


@pytest.mark.parametrize("C", [32, 64, 128, 256])
def test_prep_split_gn_is_bit_identical_to_finalize_then_split(C):
    """vt_prep_split_gn derives the GroupNorm affine in-kernel; the planes must equal vt_gn_finalize -> vt_prep_split exactly,
    also when the normalised tensor is a channel slice of a wider one (strided statistics)."""
    from vistracker_b200 import ops
    g = torch.Generator().manual_seed(100 + C)
    n, H, W = 3, 8, 16
    wide = torch.randn(n, 2 * C, H, W, generator=g) * 2 + 0.3
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    xw = nhwc(wide)
    x = xw[..., C // 2:C // 2 + C]                                  # channel slice, ld = 2C
    stats_w = chan_stats(wide).to(dev())
    stats = stats_w[:, C // 2:C // 2 + C]
    sc, sh = ops.gn_finalize(stats, gamma.to(dev()), beta.to(dev()), H * W)
    ref, _ = ops.prep_split(x, sc, sh, True, 1)
    got, ovf = ops.prep_split_gn(x, stats, gamma.to(dev()), beta.to(dev()), True, 1)
    torch.cuda.synchronize()
    assert int(ovf.item()) == 0
    assert torch.equal(ref.view(torch.int16), got.view(torch.int16))
