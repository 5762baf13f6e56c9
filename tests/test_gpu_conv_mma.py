"""The tcgen05 convolution against fp64 F.conv2d and against the CUDA-core kernel on the same device."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
from test_gpu_ops import chan_stats, dev, nchw, nhwc

pytestmark = pytest.mark.gpu

CASES = [  # ks, cin, cout, n, H, W
    (1, 64, 32, 1, 32, 32),
    (3, 64, 64, 2, 32, 32),
    (3, 32, 32, 1, 64, 64),
    (3, 128, 64, 2, 64, 64),
    (3, 256, 128, 1, 128, 128),
    (1, 256, 256, 1, 128, 128),
    (3, 64, 64, 1, 8, 256),
    (3, 128, 128, 2, 4, 128),
    (1, 128, 256, 3, 16, 16),
    # more tiles than SMs: the persistent kernel walks 2-5 tiles per CTA (both accumulator sets, phase bits flip twice)
    (1, 64, 64, 4, 128, 128),
    (3, 64, 128, 2, 128, 128),
    (3, 64, 32, 5, 128, 128),
    (1, 256, 256, 3, 128, 128),
]


@pytest.fixture(params=["persist", "persist_nostrip", "persist_e8", "plain", "strip", "pair"])
def variant(request, monkeypatch):
    """persist (the default path) = one CTA per SM walking tiles with double-buffered TMEM accumulators and the merged N = 2 BN MMA,
    A strips for 3x3 on maps >= 128 wide (VT_CONV_PERSIST=4: for every BN; the default uses them for BN = 128 only) and a resident
    resident weight panel for 1x1; persist_nostrip (VT_CONV_PERSIST=3) = the same without strips; persist_e8 (=6) = 1x1 with a second
    epilogue warp group instead of the resident panel;
    plain (VT_CONV_PERSIST=0) = one tile per CTA; VT_CONV_STRIP=1 = one A strip serves the three dx taps, =2 = strip + two images per
    CTA sharing the weight tiles.  The strip kernels serve 3x3 convolutions on maps at least 128 wide ('pair' needs an even image count)."""
    monkeypatch.setenv("VT_CONV_STRIP", {"persist": "0", "persist_nostrip": "0", "persist_e8": "0", "plain": "0", "strip": "1", "pair": "2"}[request.param])
    monkeypatch.setenv("VT_CONV_PERSIST", {"persist": "4", "persist_nostrip": "3", "persist_e8": "6"}.get(request.param, "0"))
    return request.param


@pytest.mark.parametrize("ks,cin,cout,n,H,W", CASES)
def test_conv_mma_matches_fp64(ks, cin, cout, n, H, W, variant):
    from vistracker_b200 import ops
    if variant == "persist_e8" and ks != 1:
        pytest.skip("the second epilogue warp group only serves 1x1 convolutions")
    if variant in ("strip", "pair") and not (ks == 3 and W >= 128):
        pytest.skip("the strip kernels only serve 3x3 convolutions on maps at least 128 wide")
    if variant == "pair":
        n = 2 * n
    g = torch.Generator().manual_seed(ks * 7919 + cin * 31 + cout + H)
    x = torch.randn(n, cin, H, W, generator=g)
    w = torch.randn(cout, cin, ks, ks, generator=g) * 0.05
    bias = torch.randn(cout, generator=g)
    res = torch.randn(n, cout, H, W, generator=g)
    gamma, beta = torch.randn(cin, generator=g), torch.randn(cin, generator=g)
    sc, sh = ops.gn_finalize(chan_stats(x).to(dev()), gamma.to(dev()), beta.to(dev()), H * W)
    planes, ovf = ops.prep_split(nhwc(x), sc, sh, True, ks // 2)
    st = ops.new_stats(n, cout, dev())
    out = ops.conv_mma(planes, H, W, ks // 2, w, bias.to(dev()), nhwc(res), stats=st)
    torch.cuda.synchronize()
    assert int(ovf.item()) == 0
    a = F.relu(F.group_norm(x.double(), 32, gamma.double(), beta.double(), 1e-5))
    ref = F.conv2d(a, w.double(), bias.double(), padding=ks // 2) + res.double()
    assert rel_err(nchw(out), ref) < 5e-6
    assert rel_err(st.cpu(), chan_stats(ref)) < 1e-5
    ffma = ops.conv_ffma(nhwc(x), sc, sh, True, w, bias.to(dev()), nhwc(res))
    assert rel_err(out.cpu(), ffma.cpu()) < 5e-6


def test_conv_mma_slice_and_inplace_residual():
    from vistracker_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 64, 32, 32, generator=g)
    w = torch.randn(64, 64, 3, 3, generator=g) * 0.05
    planes, _ = ops.prep_split(nhwc(x), None, None, False, 1)
    wide = torch.randn(1, 32, 32, 128, generator=g).to(dev())
    keep = wide.clone()
    ops.conv_mma(planes, 32, 32, 1, w, None, wide[..., 64:], out=wide[..., 64:])
    ref = F.conv2d(x.double(), w.double(), padding=1) + keep[..., 64:].permute(0, 3, 1, 2).double().cpu()
    assert rel_err(nchw(wide[..., 64:]), ref) < 5e-6
    assert torch.equal(wide[..., :64], keep[..., :64])


def test_conv_mma_rejects_untileable_shapes():
    from vistracker_b200 import ops
    x = torch.zeros(1, 12, 12, 64, device=dev())
    planes, _ = ops.prep_split(x, None, None, False, 1)
    with pytest.raises(RuntimeError, match="vt_conv_mma"):
        ops.conv_mma(planes, 12, 12, 1, torch.zeros(64, 64, 3, 3))


@pytest.mark.parametrize("persist", ["1", "0"])
def test_conv_mma_dual_output_fuses_the_identity_residual(persist, monkeypatch):
    """vt_conv_mma_dual: `out` keeps the raw conv slice (+ its statistics), `out2 = out + res2` is the ConvBlock output slice."""
    monkeypatch.setenv("VT_CONV_PERSIST", persist)
    from vistracker_b200 import _lib, ops
    from vistracker_b200.weights import pack_conv
    g = torch.Generator().manual_seed(11)
    n, H, W, cin, cout = 6, 64, 64, 128, 64            # 192 tiles: more than one per CTA on the persistent path
    x = torch.randn(n, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) * 0.05
    res2 = torch.randn(n, cout, H, W, generator=g)
    planes, _ = ops.prep_split(nhwc(x), None, None, False, 1)
    pk = pack_conv(w.to(dev()))
    out, out2 = torch.empty(n, H, W, cout, device=dev()), torch.empty(n, H, W, cout, device=dev())
    st, st2 = ops.new_stats(n, cout, dev()), ops.new_stats(n, cout, dev())
    r2 = nhwc(res2)
    P = _lib.ptr
    _lib.call("vt_conv_mma_dual", P(planes[0]), P(planes[1]), n, H, W, pk["cin_pad"], 1, 3, P(pk["hi"]), P(pk["lo"]), cout, None, None, 0,
              P(out), cout, P(st), cout, P(out2), cout, P(r2), cout, P(st2), cout, _lib.stream_ptr())
    ref = F.conv2d(x.double(), w.double(), padding=1)
    assert rel_err(nchw(out), ref) < 5e-6 and rel_err(st.cpu(), chan_stats(ref)) < 1e-5
    assert rel_err(nchw(out2), ref + res2.double()) < 5e-6 and rel_err(st2.cpu(), chan_stats(ref + res2.double())) < 1e-5


@pytest.mark.parametrize("fuse", ["1", "0"])
def test_encoder_plan_with_and_without_fused_residual_matches_oracle(fuse, monkeypatch):
    """VT_FUSE_RESIDUAL=1 (default) routes equal-width ConvBlocks through the dual-output epilogue, =0 through vt_add; results must
    not change."""
    monkeypatch.setenv("VT_FUSE_RESIDUAL", fuse)
    from oracle import sifnet_ref as R
    from vistracker_b200 import CHORETriplaneVisibility, default_options, resolve_dims
    from vistracker_b200.synth import synthetic_frames, synthetic_state_dict
    dims = resolve_dims(default_options())
    sd = synthetic_state_dict(dims, seed=0)
    net = CHORETriplaneVisibility(default_options(), device="cuda:0").eval()
    net.load_state_dict(sd)
    assert net._rgb.fuse_residual == (fuse == "1")
    images, *_ = synthetic_frames(1, size=512, seed=77, n_points=4)
    net.filter(images.cuda())
    with torch.no_grad():
        maps = R.sif_filter(sd, images)
    assert rel_err(net.im_feat_list[0].cpu(), maps["im_feat"]) < 1e-4
    assert rel_err(net.triplane_feat_list[2][0].cpu(), maps["tri_feat"][2]) < 1e-4
