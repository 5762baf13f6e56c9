"""Rasteriser kernels on the B200 against the numpy restatement of neural_renderer's algorithm (oracle/raster_ref.py; parity
UNPINNED: the third-party package itself is not available, see the oracle's header)."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import raster_ref as R

pytestmark = pytest.mark.gpu


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _mesh(seed, n=10):
    """A closed triangulated surface: convex hull of random points on an ellipsoid."""
    from scipy.spatial import ConvexHull
    rng = np.random.Generator(np.random.PCG64(seed))
    p = rng.standard_normal((n, 3)); p /= np.linalg.norm(p, axis=1, keepdims=True)
    p *= np.array([0.5, 0.35, 0.4])
    hull = ConvexHull(p)
    return p.astype(np.float32), hull.simplices.astype(np.int64)


def _pose(verts, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.uniform(-0.6, 0.6, 3)
    Rx = np.array([[1, 0, 0], [0, np.cos(a[0]), -np.sin(a[0])], [0, np.sin(a[0]), np.cos(a[0])]])
    Ry = np.array([[np.cos(a[1]), 0, np.sin(a[1])], [0, 1, 0], [-np.sin(a[1]), 0, np.cos(a[1])]])
    return (verts @ (Rx @ Ry).T + np.array([rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1), 2.4])).astype(np.float32)


def test_silhouette_forward_and_pseudo_gradient_match_the_restatement():
    _need_gpu()
    from vistracker_b200.render import SilhouetteRenderer
    size = 40
    verts0, faces = _mesh(0)
    K4 = np.array([1.9, 1.9, 0.5, 0.5])
    K = torch.tensor([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], dtype=torch.float32)
    batch = [_pose(verts0, s) for s in (1, 2)]
    ref_img = [R.rasterize(R.faces_of(R.project(_pose(verts0, s + 10).astype(np.float64), K4), faces), size)[1] for s in (1, 2)]
    rend = SilhouetteRenderer(faces, size, K[None].repeat(2, 1, 1), "cuda:0")
    v = torch.from_numpy(np.stack(batch)).cuda().requires_grad_(True)
    img = rend(v)
    keep = torch.ones(2, size, size, device="cuda"); keep[:, :6] = 0           # an occluder strip
    target = torch.from_numpy(np.stack(ref_img)).float().cuda()
    loss = ((keep * img - target) ** 2).sum()
    loss.backward()
    for b in range(2):
        vb = batch[b].astype(np.float64)
        fv = R.faces_of(R.project(vb, K4), faces)
        idx, alpha, _ = R.rasterize(fv, size)
        assert 30 < alpha.sum() < size * size / 2
        assert np.array_equal(img[b].detach().cpu().numpy(), alpha.astype(np.float32)), "coverage differs"
        g_alpha = 2 * (keep[b].cpu().numpy() * alpha - ref_img[b]) * keep[b].cpu().numpy()
        g_faces = R.backward_faces(fv, idx, alpha, g_alpha, size)
        g_ref = R.backward_verts(g_faces, vb, faces, K4)
        assert np.abs(g_ref).max() > 0
        assert rel_err(v.grad[b].cpu(), g_ref) < 2e-3


def test_triplane_occupancy_matches_the_restatement():
    _need_gpu()
    from vistracker_b200.render import TriplaneNrRenderer
    verts, faces = _mesh(4, n=14)
    verts = verts * 1.3
    tr = TriplaneNrRenderer(image_size=24, device="cuda:0")
    masks = tr.render_3views(faces, torch.from_numpy(verts)[None])
    assert masks.shape == (1, 3, 24, 24) and masks.dtype == torch.uint8
    for vi, view in enumerate(("right", "back", "top")):
        local = TriplaneNrRenderer.transform_view(torch.from_numpy(verts), view).numpy().astype(np.float64)
        _, alpha, depth = R.rasterize(R.faces_of(local, faces), 48)
        ref = (depth.reshape(24, 2, 24, 2).mean((1, 3)) < R.FAR)          # average-pooled depth < far  (anti_aliasing=True)
        assert np.array_equal(masks[0, vi].cpu().numpy().astype(bool), ref), view
        assert 20 < ref.sum() < 24 * 24


def test_axis_aligned_square_is_rendered_with_flipped_rows():
    _need_gpu()
    from vistracker_b200.render import SilhouetteRenderer
    # orthographic: a quad covering x in [-0.5, 0.25], y in [0.0, 0.75] (NDC, y up) at depth 5
    verts = torch.tensor([[[-0.5, 0.0, 5.0], [0.25, 0.0, 5.0], [0.25, 0.75, 5.0], [-0.5, 0.75, 5.0]]], device="cuda")
    faces = np.array([[0, 1, 2], [0, 2, 3]])
    img = SilhouetteRenderer(faces, 16, None, "cuda:0")(verts)[0].cpu().numpy()
    xs = (2 * np.arange(16) + 1 - 16) / 16
    exp = np.zeros((16, 16), np.float32)
    for r in range(16):
        for c in range(16):
            yp, xp = xs[15 - r], xs[c]                      # image row 0 is the top (y = +1)
            exp[r, c] = float(-0.5 <= xp <= 0.25 and 0.0 <= yp <= 0.75)
    assert np.array_equal(img, exp)


def test_chunk_culling_does_not_change_the_image(monkeypatch):
    """vt_raster_fwd with the bounding-box workspace (tiles skip the 256-face chunks that miss them) against the exhaustive scan: depth,
    coverage are bit-identical, the backward pass (which consumes the face-index map) agrees to rounding; several chunks, faces straddling tiles, B = 3."""
    _need_gpu()
    from vistracker_b200.render import SilhouetteRenderer
    verts0, faces = _mesh(7, n=500)                                                # ~1000 faces -> 2000 with the reversed copies: 8 chunks
    assert faces.shape[0] > 600
    batch = torch.from_numpy(np.stack([_pose(verts0, s) for s in (1, 2, 3)])).cuda()
    K = torch.tensor([[1.9, 0, 0.5], [0, 1.9, 0.5], [0, 0, 1]], dtype=torch.float32)[None].repeat(3, 1, 1)
    out = {}
    for cull in ("1", "0"):
        monkeypatch.setenv("VT_RASTER_CULL", cull)
        rend = SilhouetteRenderer(faces, 200, K, "cuda:0")
        v = batch.clone().requires_grad_(True)
        img = rend(v)
        (img * torch.linspace(-1, 1, 200, device="cuda")[None, None]).sum().backward()
        ortho = SilhouetteRenderer(faces, 136, None, "cuda:0").render_depth(batch - torch.tensor([0, 0, 2.0], device="cuda"))
        out[cull] = (img.detach().clone(), v.grad.clone(), ortho.clone())
    assert 2000 < float(out["1"][0].sum()) < 3 * 200 * 200 / 2
    assert torch.equal(out["1"][0], out["0"][0]) and torch.equal(out["1"][2], out["0"][2])
    assert rel_err(out["1"][1].cpu(), out["0"][1].cpu()) < 1e-5                    # same face-index map; the vertex scatter uses atomics


def test_silloss_roi_from_masks_constructor():
    """The reference constructor signature (person masks, object masks, template, crop centres): ROI set-up on the device equals the CPU
    run of the same function, and the loss renders into the ROI camera."""
    _need_gpu()
    from vistracker_b200.render import SilLossROI, roi_setup
    verts0, faces = _mesh(3, n=40)
    om = torch.zeros(2, 512, 512); om[:, 200:300, 260:340] = 1.0
    pm = torch.zeros(2, 512, 512); pm[:, 150:400, 180:280] = 1.0
    cc = torch.tensor([[1024.0, 768.0], [1000.0, 800.0]])
    sil = SilLossROI.from_masks(pm.cuda(), om.cuda(), verts0 * 0.5, faces, cc.cuda(), device="cuda:0")
    keep, ref, K = roi_setup(pm, om, cc)
    assert torch.equal(sil.keep_mask.cpu(), keep) and torch.equal(sil.image_ref.cpu(), ref)
    R = torch.eye(3, device="cuda")[None].repeat(2, 1, 1).requires_grad_(True)
    t = torch.tensor([[0.05, 0.0, 2.3], [0.0, 0.05, 2.3]], device="cuda", requires_grad=True)
    out = sil.forward(R, t, torch.ones(2, device="cuda"))
    loss = out[0]["mask"] if isinstance(out, tuple) else out["mask"]
    loss.sum().backward()
    assert bool(torch.isfinite(loss).all()) and t.grad is not None and bool(torch.isfinite(t.grad).all())


def test_backward_walk_skipping_does_not_change_the_gradient():
    """vt_raster_bwd_ws (prefix counts of the contributing background pixels: walks that cannot contribute are skipped) against the plain
    vt_raster_bwd walk: the same gradient (to the rounding of the atomic vertex scatter), on a target that is offset from the render (a band of contributing pixels), with an
    occluder strip, and on a target equal to the render (nothing contributes)."""
    _need_gpu()
    from vistracker_b200.render import SilhouetteRenderer
    size = 96
    verts0, faces = _mesh(4, n=40)
    K4 = np.array([1.9, 1.9, 0.5, 0.5])
    K = torch.tensor([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], dtype=torch.float32)
    batch = np.stack([_pose(verts0, s) for s in (1, 2, 3)])
    rend = SilhouetteRenderer(faces, size, K[None].repeat(3, 1, 1), "cuda:0")
    with torch.no_grad():
        shifted = rend(torch.from_numpy(np.stack([_pose(verts0, s + 10) for s in (1, 2, 3)])).cuda())
        same = rend(torch.from_numpy(batch).cuda())
    keep = torch.ones(3, size, size, device="cuda"); keep[:, :, :9] = 0
    for ti, target in enumerate((shifted, same)):
        grads = []
        for skip in (False, True):
            rend.skip_walks = skip
            v = torch.from_numpy(batch).cuda().requires_grad_(True)
            ((keep * rend(v) - target) ** 2).sum().backward()
            grads.append(v.grad.clone())
        # the per-face gradients are identical; the scatter to the vertices adds them with float atomics (order varies run to run)
        if ti == 0:
            assert float(grads[0].abs().max()) > 0 and rel_err(grads[1].cpu(), grads[0].cpu()) < 1e-6
        else:
            assert float((grads[1] - grads[0]).abs().max()) <= 1e-6 * max(float(grads[0].abs().max()), 1e-30)
    rend.skip_walks = True


def test_many_layers_over_one_pixel_keep_the_nearest_face_in_face_order():
    """Up to 13 faces stacked over the same pixels, nearest one in the middle of the face list and two faces at EXACTLY the same depth: the forward
    kernel queues the faces that cover a pixel (four per queue) and evaluates their depths later -- the queue must flush in the middle of the
    face loop without losing the order ('nearest wins, first one on ties').  Depth image and face-index map against the restatement."""
    _need_gpu()
    from vistracker_b200 import _lib
    from vistracker_b200.render import SilhouetteRenderer, _cull_ws
    rng = np.random.Generator(np.random.PCG64(5))
    S, L = 48, 13
    depth_of = [9.0, 7.5, 8.0, 6.0, 6.5, 3.0, 5.0, 3.0, 4.0, 2.5, 2.5, 7.0, 2.75]          # nearest: faces 9 and 10 (a tie: 9 wins), 5 and 7 tie behind them
    verts, faces = [], []
    for l in range(L):
        c = rng.uniform(-0.25, 0.25, 2)
        r = rng.uniform(0.45, 0.8)
        a0 = rng.uniform(0, 2 * np.pi)
        for k in range(3):
            verts.append([c[0] + r * np.cos(a0 + 2.1 * k), c[1] + r * np.sin(a0 + 2.1 * k), depth_of[l]])
        faces.append([3 * l, 3 * l + 1, 3 * l + 2])
    verts, faces = np.asarray(verts, np.float32), np.asarray(faces, np.int64)
    rend = SilhouetteRenderer(faces, S, None, "cuda:0")                           # orthographic: (x, y, z) are NDC
    v = torch.from_numpy(verts)[None].cuda().contiguous()
    F_ = faces.shape[0]
    faces_ndc = torch.empty(1, 2 * F_, 9, device="cuda")
    fidx = torch.empty(1, S, S, dtype=torch.int32, device="cuda")
    depth = torch.empty(1, S, S, device="cuda")
    cull = _cull_ws(1, F_, v.device)
    _lib.call("vt_raster_fwd", _lib.ptr(v), _lib.ptr(rend.faces), 1, v.shape[1], F_, rend.mode, _lib.ptr(rend.K4), S, _lib.ptr(faces_ndc), _lib.ptr(fidx),
              None, _lib.ptr(depth), _lib.ptr(cull), _lib.stream_ptr())
    torch.cuda.synchronize()
    idx_ref, alpha_ref, depth_ref = R.rasterize_fast(R.faces_of(verts.astype(np.float64), faces), S)
    fv = R.faces_of(verts.astype(np.float64), faces)
    per_face = np.stack([R.rasterize_fast(q[None], S)[2][::-1] for q in fv])      # [2F, S, S] depth of every face alone, y-up like the index map
    n_cover = (per_face < R.FAR).sum(0)
    assert n_cover.max() >= 9                                                     # more than two queues' worth of faces over some pixels
    d = depth[0].cpu().numpy()
    assert np.array_equal(d < R.FAR, depth_ref < R.FAR)
    assert np.abs(d - depth_ref).max() < 1e-5
    # the index map: exact wherever the nearest face is nearest by a margin; where two faces sit at the same depth (9 / 10 and 5 / 7 by
    # construction) fp32 rounding of the weights may order them either way, so there the winner only has to be one of the tied faces
    two = np.sort(per_face, 0)[:2]
    clear = (two[1] - two[0]) > 1e-4
    got = fidx[0].cpu().numpy()
    assert clear.sum() > 200 and (~clear & (n_cover > 1)).sum() > 20
    assert np.array_equal(got[clear], idx_ref[clear]), f"{(got[clear] != idx_ref[clear]).sum()} pixels with a clear nearest face differ"
    amb = ~clear & (n_cover > 0)
    yy, xx = np.nonzero(amb)
    assert all(abs(per_face[got[y, x], y, x] - two[0][y, x]) < 1e-5 for y, x in zip(yy, xx))
