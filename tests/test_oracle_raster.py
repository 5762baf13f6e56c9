"""The vectorised rasteriser of the oracle (used by the optimize_smpl_object loop golden) against its plain-loop twin."""
import numpy as np

from oracle import raster_ref as R


def _mesh(seed, n):
    from scipy.spatial import ConvexHull
    rng = np.random.default_rng(seed)
    p = rng.standard_normal((n, 3)); p /= np.linalg.norm(p, axis=1, keepdims=True)
    p = p * [0.3, 0.25, 0.2] + [rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1), 2.3]
    return p, ConvexHull(p).simplices


def test_vectorised_rasteriser_equals_the_loops():
    for seed, n, size in ((0, 30, 24), (1, 12, 31)):
        v, f = _mesh(seed, n)
        fv = R.faces_of(R.project(v, (2.0, 2.0, 0.5, 0.5)), f)
        idx_a, alpha_a, depth_a = R.rasterize(fv, size)
        idx_b, alpha_b, depth_b = R.rasterize_fast(fv, size)
        assert alpha_a.sum() > 20
        assert np.array_equal(idx_a, idx_b) and np.array_equal(alpha_a, alpha_b)
        assert np.abs(depth_a - depth_b).max() < 1e-12


def test_brute_force_coverage_agrees_with_the_restatement():
    """An independent float64 statement of what a silhouette is -- a pixel centre is covered iff it lies inside (or on the border of) the
    2-D projection of some triangle whose interpolated depth is in (near, far) -- sharing no code with rasterize(): signed-area
    point-in-triangle test on both windings.  Pixels whose centre is within 1e-9 of an edge are excluded (tie rules differ)."""
    v, f = _mesh(3, 20)
    size = 32
    ndc = R.project(v, (2.2, 2.2, 0.5, 0.5))
    _, alpha, _ = R.rasterize_fast(R.faces_of(ndc, f), size)
    cover = np.zeros((size, size), bool); near_edge = np.zeros((size, size), bool)
    ys, xs = np.mgrid[0:size, 0:size]
    px, py = (2 * xs + 1 - size) / size, (2 * ys + 1 - size) / size
    for tri in f:
        a, b, c = ndc[tri[0], :2], ndc[tri[1], :2], ndc[tri[2], :2]
        cross = lambda p, q: (q[0] - p[0]) * (py - p[1]) - (q[1] - p[1]) * (px - p[0])
        e0, e1, e2 = cross(a, b), cross(b, c), cross(c, a)
        inside = ((e0 >= 0) & (e1 >= 0) & (e2 >= 0)) | ((e0 <= 0) & (e1 <= 0) & (e2 <= 0))
        cover |= inside
        near_edge |= (np.minimum(np.minimum(np.abs(e0), np.abs(e1)), np.abs(e2)) < 1e-9)
    cover = cover[::-1]; near_edge = near_edge[::-1]                   # image rows are flipped (y up in NDC)
    ok = ~near_edge
    assert cover.sum() > 30 and np.array_equal(cover[ok], alpha[ok] > 0)
