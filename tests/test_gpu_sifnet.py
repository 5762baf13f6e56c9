"""SIF-Net filter + query on the B200 through the reference-shaped API, against the reference golden vectors
(tests/golden, produced by the unmodified reference) and against the CPU oracle on fresh seeded inputs.

Tolerance: north_star asks for 1e-4 relative; ``rel_err`` is max|a-b| / max|b| per tensor."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import sifnet_ref as R
from vistracker_b200 import CHORETriplaneVisibility, default_options, resolve_dims
from vistracker_b200.synth import synthetic_frames, synthetic_state_dict

pytestmark = pytest.mark.gpu
TOL = 1e-4
DIMS = resolve_dims(default_options())
CAM = (DIMS.fx_px, DIMS.fy_px, DIMS.cx_px, DIMS.cy_px, DIMS.crop_size)


@pytest.fixture(scope="module")
def net():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    m = CHORETriplaneVisibility(default_options(), device="cuda:0").eval()
    m.load_state_dict(synthetic_state_dict(DIMS, seed=0))
    return m


def _maps_nchw(net):
    return {"im_feat": net.im_feat_list[0], "tmpx": net.tmpx,
            **{f"tri_feat{v}": net.triplane_feat_list[v][0] for v in range(3)},
            **{f"tri_tmpx{v}": net.triplane_tmpx[v] for v in range(3)}}


def _check_heads(net, g):
    for name, o in zip(("df", "pca", "parts", "centers", "vis"), net.get_preds()):
        assert tuple(o.shape) == g[name].shape, name
        assert rel_err(o.cpu(), g[name]) < TOL, name


def test_small_frames_match_reference_golden(net, golden):
    """64x64 frames: feature maps are too small for 128-pixel tensor-core tiles -> exercises the CUDA-core conv path."""
    g = golden("sifnet_small.npz")
    images, points, crop, body = synthetic_frames(2, size=64, seed=11, n_points=301, jitter=True)
    net.filter(images.cuda())
    for k, v in _maps_nchw(net).items():
        assert rel_err(v.cpu(), g[k]) < TOL, k
    net.query(points.cuda(), crop_center=crop.cuda(), body_center=body.cuda())
    _check_heads(net, g)
    assert rel_err(net.points_xy.cpu(), g["xy"]) < 1e-6
    feat, xy = net.query_features(points.cuda(), crop.cuda(), body_center=body.cuda())
    assert rel_err(feat.cpu(), g["features"]) < TOL
    out_of_img = (np.abs(g["xy"]) > 1).any(1)
    assert out_of_img.any()
    assert (net.get_preds()[0].cpu().numpy().transpose(0, 2, 1)[out_of_img] == 5.0).all()


def test_config1_matches_reference_golden(net, golden):
    """BASELINE config 1: one 512x512 frame + 2000 points; tcgen05 convolutions everywhere past the stem."""
    g = golden("sifnet_c1.npz")
    images, points, crop, body = synthetic_frames(1, size=512, seed=0, n_points=2000)
    net.filter(images.cuda())
    for k, v in _maps_nchw(net).items():
        assert rel_err(v[:, :, ::8, ::8].cpu(), g[k]) < TOL, k
    net.query(points.cuda(), crop_center=crop.cuda(), body_center=body.cuda())
    _check_heads(net, g)


def test_batch_matches_oracle(net):
    """B=2 at full resolution with jittered crops / body centres against the CPU oracle (fresh seed, no golden)."""
    sd = synthetic_state_dict(DIMS, seed=0)
    images, points, crop, body = synthetic_frames(2, size=512, seed=21, n_points=1500, jitter=True)
    with torch.no_grad():
        maps = R.sif_filter(sd, images)
        ref = R.sif_query(sd, maps, points, crop, body, CAM)
    net.filter(images.cuda())
    assert rel_err(net.im_feat_list[0].cpu(), maps["im_feat"]) < TOL
    assert rel_err(net.tmpx.cpu(), maps["tmpx"]) < TOL
    for v in range(3):
        assert rel_err(net.triplane_feat_list[v][0].cpu(), maps["tri_feat"][v]) < TOL
        assert rel_err(net.triplane_tmpx[v].cpu(), maps["tri_tmpx"][v]) < TOL
    net.query(points.cuda(), crop_center=crop.cuda(), body_center=body.cuda())
    for o, r in zip(net.get_preds(), ref):
        assert rel_err(o.cpu(), r) < TOL


def test_config2_shape_matches_oracle(net):
    """BASELINE config 2 -- the shape bench.py's C2 numbers are quoted on: 8 frames of 512 x 512, 10 000 query points per frame.  Frames 0 and 5
    of the batch are compared with the CPU oracle head by head, in both tolerance forms: max|a - b| / max|b| (north_star's 1e-4) and the
    per-element form of SURVEY.md 7 (denominator max(|b|, 1e-3 max|b|))."""
    from conftest import elementwise_err
    sd = synthetic_state_dict(DIMS, seed=0)
    images, points, crop, body = synthetic_frames(8, size=512, seed=33, n_points=10000, jitter=True)
    net.filter(images.cuda())
    net.query(points.cuda(), crop_center=crop.cuda(), body_center=body.cuda())
    ours = [o.cpu() for o in net.get_preds()]
    worst = {}
    for f in (0, 5):
        with torch.no_grad():
            maps = R.sif_filter(sd, images[f:f + 1])
            ref = R.sif_query(sd, maps, points[f:f + 1], crop[f:f + 1], body[f:f + 1], CAM)
        for name, o, r in zip(("df", "pca", "parts", "centers", "vis"), ours, ref):
            e1, e2 = rel_err(o[f:f + 1], r), elementwise_err(o[f:f + 1], r)
            worst[name] = (max(worst.get(name, (0, 0))[0], e1), max(worst.get(name, (0, 0))[1], e2))
    print("C2 shape, worst (max-rel, element-wise) per head:", {k: (f"{a:.2e}", f"{b:.2e}") for k, (a, b) in worst.items()})
    for name, (e1, e2) in worst.items():
        assert e1 < TOL, (name, e1)
        assert e2 < 1e-4, (name, e2)                       # north_star's 1e-4, per element


def test_query_edge_cases(net):
    images, points, crop, body = synthetic_frames(1, size=64, seed=3, n_points=33)
    net.filter(images.cuda())
    # N not a multiple of the 32-point tile, N = 1, and points behind / on the camera plane
    for n in (33, 1):
        net.query(points[:, :n].cuda(), crop_center=crop.cuda(), body_center=body.cuda())
        assert net.get_preds()[0].shape == (1, 2, n)
    bad = points[:, :4].clone(); bad[0, 0, 2] = 0.0; bad[0, 1, 2] = -1.0
    net.query(bad.cuda(), crop_center=crop.cuda(), body_center=body.cuda())
    assert torch.isfinite(net.get_preds()[2]).all()
    with pytest.raises(ValueError):
        net.query(points.repeat(2, 1, 1).cuda(), crop_center=crop.cuda(), body_center=body.cuda())


def _grad_wrt_points(net, points, crop, body, idx):
    pts = points.cuda().requires_grad_(True)
    net.query(pts, crop_center=crop.cuda(), body_center=body.cuda())
    df = net.get_preds()[0]
    torch.clamp(df[:, idx], max=2.0).sum().backward()         # recon/gen/generator.py:88-90
    return pts.grad.cpu()


def test_query_gradient_matches_reference_golden(net, golden):
    """d(sum clamp(df, 2))/d(points): the quantity Generator.approx_surface back-propagates 10x per round."""
    g = golden("sifnet_small.npz")
    images, points, crop, body = synthetic_frames(2, size=64, seed=11, n_points=301, jitter=True)
    net.filter(images.cuda())
    for name, idx in (("grad_h", 0), ("grad_o", 1)):
        got = _grad_wrt_points(net, points, crop, body, idx)
        assert rel_err(got, g[name]) < TOL, name
        out_of_img = (np.abs(g["xy"]) > 1).any(1)
        assert float(got.numpy()[out_of_img].__abs__().max()) == 0.0      # df is a constant 5.0 there
    g = golden("sifnet_c1.npz")
    images, points, crop, body = synthetic_frames(1, size=512, seed=0, n_points=2000)
    net.filter(images.cuda())
    for name, idx in (("grad_h", 0), ("grad_o", 1)):
        assert rel_err(_grad_wrt_points(net, points, crop, body, idx), g[name]) < TOL, name


def test_query_gradient_all_heads_matches_oracle(net):
    """Random cotangents on all 29 outputs (parts / centres / visibility heads are used by the fitters' losses)."""
    sd = synthetic_state_dict(DIMS, seed=0)
    images, points, crop, body = synthetic_frames(2, size=64, seed=31, n_points=77, jitter=True)
    net.filter(images.cuda())
    cot = [torch.randn(2, c, 77, generator=torch.Generator().manual_seed(40 + c)) for c in (2, 9, 14, 3, 1)]
    with torch.no_grad():
        maps = R.sif_filter(sd, images)
    p_ref = points.clone().requires_grad_(True)
    outs = R.sif_query(sd, maps, p_ref, crop, body, CAM)
    sum((o.reshape(2, -1, 77) * c).sum() for o, c in zip(outs, cot)).backward()
    pts = points.cuda().requires_grad_(True)
    net.query(pts, crop_center=crop.cuda(), body_center=body.cuda())
    sum((o.reshape(2, -1, 77) * c.cuda()).sum() for o, c in zip(net.get_preds(), cot)).backward()
    assert rel_err(pts.grad.cpu(), p_ref.grad) < TOL


def test_tensor_core_and_cuda_core_decoders_agree(net):
    """query_fwd_tc_kernel (tcgen05, default) against query_fwd_kernel (fp32 FFMA) on the same maps; N not a multiple of 128."""
    images, points, crop, body = synthetic_frames(2, size=64, seed=41, n_points=777, jitter=True)
    net.filter(images.cuda())
    assert not net.query_on_cuda_cores
    out_tc, xy_tc = net._query_raw(points.cuda(), crop.cuda(), body.cuda(), want_xy=True)
    net.query_on_cuda_cores = True
    try:
        out_cc, xy_cc = net._query_raw(points.cuda(), crop.cuda(), body.cuda(), want_xy=True)
    finally:
        net.query_on_cuda_cores = False
    net.check()
    assert torch.equal(xy_tc, xy_cc)
    for lo, hi in ((0, 2), (2, 11), (11, 25), (25, 28), (28, 29)):
        assert rel_err(out_tc[:, lo:hi].cpu(), out_cc[:, lo:hi].cpu()) < 2e-5


@pytest.mark.parametrize("head_mask,scale", [(31, 1.0), (1, 1.0), (5, 3e-8), (16, 1e3), (10, 1.0)])
def test_tensor_core_and_cuda_core_backward_agree(net, head_mask, scale):
    """query_bwd_tc_kernel (tcgen05, default) against query_bwd_kernel (fp32 FFMA): head subsets, N not a multiple of 128 and
    cotangents as small as a mean over 1e7 elements would make them (the per-point power-of-two renormalisation must keep them
    out of the fp16 underflow range)."""
    B, N = 2, 333
    images, points, crop, body = synthetic_frames(B, size=64, seed=51, n_points=N, jitter=True)
    net.filter(images.cuda())
    g = torch.randn(B, 29, N, generator=torch.Generator().manual_seed(head_mask)) * scale
    for h, (lo, hi) in enumerate(((0, 2), (2, 11), (11, 25), (25, 28), (28, 29))):
        if not (head_mask >> h) & 1:
            g[:, lo:hi] = 0
    args = (points.cuda(), crop.cuda(), body.cuda(), g.cuda())
    assert not net.query_on_cuda_cores
    got = net._query_backward(*args, head_mask=head_mask)
    net.query_on_cuda_cores = True
    try:
        ref = net._query_backward(*args, head_mask=head_mask)
    finally:
        net.query_on_cuda_cores = False
    net.check()
    assert float(ref.abs().max()) > 0
    assert rel_err(got.cpu(), ref.cpu()) < 2e-5


def test_tensor_core_projection_step_matches_cuda_core_step(net):
    """vt_query_project_step_tc against vt_query_project_step (one Generator.approx_surface step, both distance channels)."""
    from vistracker_b200.generator import GeneratorTriplaneVis
    B, N = 2, 500
    images, points, crop, body = synthetic_frames(B, size=64, seed=61, n_points=N, jitter=True)
    net.filter(images.cuda())
    gen = GeneratorTriplaneVis(net, threshold=2.0)
    qi = {"crop_center": crop.cuda(), "body_center": body.cuda()}
    for df_idx in (0, 1):
        new_tc, out_tc = gen._project_step(points.cuda(), qi, df_idx, True)
        net.query_on_cuda_cores = True
        try:
            new_cc, out_cc = gen._project_step(points.cuda(), qi, df_idx, True)
        finally:
            net.query_on_cuda_cores = False
        net.check()
        assert rel_err(out_tc.cpu(), out_cc.cpu()) < 2e-5
        # the update direction is normalize(grad): compare where the gradient is not vanishing (flat random-init UDF regions)
        step_tc, step_cc = (new_tc - points.cuda()).cpu(), (new_cc - points.cuda()).cpu()
        err = (step_tc - step_cc).abs().max(-1).values / step_cc.abs().max()
        assert float((err < 1e-3).float().mean()) > 0.97


def test_graph_replay_matches_eager_filter(net):
    """filter() replays a CUDA graph captured on the first call of a shape; a later call with other images must give what an eager
    pass gives for THOSE images (statistics are summed with fp64 atomics, so equality is to fp32 rounding through ~100 layers, not bitwise)."""
    assert net.use_graph
    a, *_ = synthetic_frames(2, size=64, seed=71, n_points=4)
    b, *_ = synthetic_frames(2, size=64, seed=72, n_points=4)
    net.filter(a.cuda())                       # capture (or reuse) on images a
    net.filter(b.cuda())                       # replay on images b
    got = [t.clone() for t in net._maps]
    net.use_graph = False
    try:
        net.filter(b.cuda())
        ref = [t.clone() for t in net._maps]
    finally:
        net.use_graph = True
    for g, r in zip(got, ref):
        assert rel_err(g.cpu(), r.cpu()) < 1e-5
    assert rel_err(got[0].cpu(), net.im_feat_list[0].permute(0, 2, 3, 1).cpu()) < 1e-5


@pytest.mark.parametrize("df_channel,clamp_max,with_parts", [(0, 0.1, True), (1, 0.8, False), (0, 2.0, True)])
def test_fused_fitting_losses_match_query_plus_autograd(net, df_channel, clamp_max, with_parts):
    """vt_query_losses_tc (values + point gradients in one launch) against query() + torch.clamp / F.cross_entropy + autograd on the
    fp32 CUDA-core kernels, with per-frame weights on the distance term as forward_step applies them."""
    B, N = 2, 419
    images, points, crop, body = synthetic_frames(B, size=64, seed=81, n_points=N, jitter=True)
    net.filter(images.cuda())
    labels = torch.randint(0, 14, (B, N), generator=torch.Generator().manual_seed(5)).cuda() if with_parts else None
    w = torch.tensor([0.3, 1.7]).cuda()

    def total(vals_df, vals_ce):
        t = (vals_df.mean(-1) * w).mean()
        return t + (0.01 * vals_ce.sum(-1).mean() if vals_ce is not None else 0.0)

    res = {}
    for cuda_cores in (False, True):
        net.query_on_cuda_cores = cuda_cores
        try:
            pts = points.cuda().requires_grad_(True)
            vals_df, vals_ce = net.query_losses(pts, crop_center=crop.cuda(), body_center=body.cuda(), df_channel=df_channel,
                                                clamp_max=clamp_max, part_labels=labels)
            total(vals_df, vals_ce).backward()
            res[cuda_cores] = (vals_df.detach().cpu(), None if vals_ce is None else vals_ce.detach().cpu(), pts.grad.cpu())
        finally:
            net.query_on_cuda_cores = False
    net.check()
    assert rel_err(res[False][0], res[True][0]) < 2e-5
    if with_parts:
        assert rel_err(res[False][1], res[True][1]) < 2e-5
    assert float(res[True][2].abs().max()) > 0
    assert rel_err(res[False][2], res[True][2]) < 5e-5


def test_query_heads_subset_matches_full_query(net):
    images, points, crop, body = synthetic_frames(2, size=64, seed=91, n_points=300, jitter=True)
    net.filter(images.cuda())
    net.query(points.cuda(), crop_center=crop.cuda(), body_center=body.cuda())
    full = dict(zip(("df", "pca", "parts", "centers", "visibility"), net.get_preds()))
    for heads in (("centers",), ("df", "parts"), ("visibility", "pca", "df", "centers", "parts")):
        sub = net.query_heads(points.cuda(), heads, crop_center=crop.cuda(), body_center=body.cuda())
        for h in heads:
            assert torch.equal(sub[h], full[h]), h


def test_maps_of_an_earlier_filter_call_stay_valid(net):
    """filter() on a second batch of the same shape must not overwrite the maps handed out for the first (graph replay copies out)."""
    a, *_ = synthetic_frames(2, size=64, seed=101, n_points=4)
    b, *_ = synthetic_frames(2, size=64, seed=102, n_points=4)
    net.filter(a.cuda())
    kept = net._maps
    snapshot = [t.clone() for t in kept]
    net.filter(b.cuda())
    for t, s in zip(kept, snapshot):
        assert torch.equal(t, s)
    assert not torch.equal(net._maps[0], kept[0])


@pytest.mark.parametrize("with_parts,also", [(False, ("centers",)), (True, ("centers", "visibility")), (False, ("pca", "parts", "centers", "visibility"))])
def test_fused_losses_with_forward_only_heads(net, with_parts, also):
    """query_losses(also=...) evaluates extra heads forward-only in the same launch (they share the feature gather with the loss heads):
    values equal query()'s, the loss terms and their gradient are unchanged by the extra heads."""
    B, N = 2, 257
    images, points, crop, body = synthetic_frames(B, size=64, seed=111, n_points=N, jitter=True)
    net.filter(images.cuda())
    labels = torch.randint(0, 14, (B, N), generator=torch.Generator().manual_seed(9)).cuda() if with_parts else None
    kw = dict(crop_center=crop.cuda(), body_center=body.cuda(), df_channel=1, clamp_max=0.8, part_labels=labels)
    p0 = points.cuda().requires_grad_(True)
    v0, c0 = net.query_losses(p0, **kw)
    (v0.sum() + (c0.sum() if c0 is not None else 0.0)).backward()
    p1 = points.cuda().requires_grad_(True)
    v1, c1, extra = net.query_losses(p1, also=also, **kw)
    (v1.sum() + (c1.sum() if c1 is not None else 0.0)).backward()
    net.check()
    assert torch.equal(v0, v1) and torch.equal(p0.grad, p1.grad)
    net.query(points.cuda(), crop_center=crop.cuda(), body_center=body.cuda())
    full = dict(zip(("df", "pca", "parts", "centers", "visibility"), net.get_preds()))
    for h in also:        # the two column halves of the last hidden layer are summed by two threads: equal to the last rounding, not bitwise
        assert rel_err(extra[h].cpu(), full[h].cpu()) < 1e-6, h


@pytest.mark.parametrize("w_df,w_ce,clamp_max", [(1.0, 1.0, 0.1), (30.0, 0.0005, 0.1), (7.5, 3.0e3, 2.0), (0.0, 1.0, 0.1)])
def test_merged_heads_launch_equals_the_weighted_sum_of_the_two_head_gradients(net, w_df, w_ce, clamp_max):
    """vt_query_losses_merged_tc (both heads' hidden gradients accumulated in tensor memory, ONE backward gather) against
    w_df * g_df + w_ce * g_ce of vt_query_losses_tc, with weights spanning the ratios the schedule of optimize_smpl reaches
    (recon_fit_behave.py:393-420: df_h 30^2 / (B V), part 0.05^2 / B and back)."""
    B, N = 2, 419
    images, points, crop, body = synthetic_frames(B, size=64, seed=121, n_points=N, jitter=True)
    net.filter(images.cuda())
    pts, cc, bc = points.cuda().contiguous(), crop.cuda(), body.cuda()
    labels = torch.randint(0, 14, (B, N), generator=torch.Generator().manual_seed(11)).cuda()
    f = lambda *s: torch.full(s, float("nan"), device="cuda")
    vd, gd, vc, gc = f(B, N), f(B, N, 3), f(B, N), f(B, N, 3)
    net.enqueue_query_losses(pts, cc, bc, 0, clamp_max, labels, vd, gd, vc, gc)
    w = torch.tensor([w_df, w_ce], device="cuda")
    vd2, vc2, gm = f(B, N), f(B, N), f(B, N, 3)
    net.enqueue_query_losses_merged(pts, cc, bc, 0, clamp_max, labels, w.data_ptr(), 0.5, w.data_ptr() + 4, 2.0, vd2, vc2, gm)
    torch.cuda.synchronize()
    net.check()
    assert torch.equal(vd, vd2) and torch.equal(vc, vc2)
    ref = (0.5 * w_df) * gd + (2.0 * w_ce) * gc
    assert float(ref.abs().max()) > 0
    assert rel_err(gm.cpu(), ref.cpu()) < 2e-5
