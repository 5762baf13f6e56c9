"""CPU restatement of the silhouette / depth rasteriser (TEST INFRASTRUCTURE).

PARITY UNPINNED.  The reference renders through ``neural_renderer`` (call sites recon/obj_pose_roi.py:87-94,192 and
render/render_triplane_nr.py:27-28,106-107), a third-party CUDA package that is neither vendored, pinned nor listed in
requirements.txt, and that cannot be installed here.  This file restates the upstream algorithm (daniilidis-group port of
Kato et al.'s Neural 3D Mesh Renderer: ``projection``, ``vertices_to_faces``, ``forward_face_index_map``,
``backward_pixel_map``) as recalled from its sources, in plain numpy loops, to cross-check the CUDA kernel.  It must be
re-validated against the real package wherever that can be installed.
"""
from __future__ import annotations

import math

import numpy as np

NEAR, FAR, EPS = 0.1, 100.0, 1e-3


def project(verts, K4):
    """nr.projection with R = I, t = 0, no distortion, orig_size = 1: verts [V,3], K4 = (fx, fy, cx, cy) -> NDC [V,3]."""
    z = verts[:, 2] + 1e-9
    px = K4[0] * (verts[:, 0] / z) + K4[2]
    py = K4[1] * (verts[:, 1] / z) + K4[3]
    return np.stack([2 * (px - 0.5), 2 * ((1 - py) - 0.5), verts[:, 2]], 1)


def faces_of(ndc, faces):
    """vertices_to_faces on the fill_back face list: [2F, 3, 3]."""
    f2 = np.concatenate([faces, faces[:, ::-1]], 0)
    return ndc[f2]


def rasterize(fv, size):
    """face index map (y-up internal layout), alpha and depth in image layout (rows flipped)."""
    idx = -np.ones((size, size), np.int64)
    depth = np.full((size, size), FAR)
    for yi in range(size):
        for xi in range(size):
            yp, xp = (2 * yi + 1 - size) / size, (2 * xi + 1 - size) / size
            best, dmin = -1, FAR
            for fn, q in enumerate(fv):
                (x0, y0, z0), (x1, y1, z1), (x2, y2, z2) = q
                if ((yp - y0) * (x1 - x0) < (xp - x0) * (y1 - y0)) or ((yp - y1) * (x2 - x1) < (xp - x1) * (y2 - y1)) or \
                        ((yp - y2) * (x0 - x2) < (xp - x2) * (y0 - y2)):
                    continue
                m = np.array([[x0, x1, x2], [y0, y1, y2], [1, 1, 1]], np.float64)
                if abs(np.linalg.det(m)) < 1e-300:
                    continue
                w = np.clip(np.linalg.inv(m) @ np.array([xp, yp, 1.0]), 0, 1)
                w = w / max(w.sum(), 1e-10)
                zp = 1.0 / (w[0] / z0 + w[1] / z1 + w[2] / z2)
                if zp <= NEAR or zp >= FAR:
                    continue
                if zp < dmin:
                    dmin, best = zp, fn
            idx[yi, xi], depth[yi, xi] = best, dmin
    alpha = (idx >= 0).astype(np.float64)
    return idx, alpha[::-1].copy(), depth[::-1].copy()


def rasterize_fast(fv, size):
    """Same result as ``rasterize`` (face index map, alpha, depth) with the face loop vectorised in numpy: per pixel row the three edge tests,
    the clipped barycentric weights and the 1 / sum(w / z) depth of every face at once; ties keep the lowest face index, as the loop does.
    ``tests/test_oracle_raster.py`` checks it against the loop version."""
    fv = np.asarray(fv, np.float64)
    x0, y0, z0 = fv[:, 0, 0], fv[:, 0, 1], fv[:, 0, 2]
    x1, y1, z1 = fv[:, 1, 0], fv[:, 1, 1], fv[:, 1, 2]
    x2, y2, z2 = fv[:, 2, 0], fv[:, 2, 1], fv[:, 2, 2]
    det = x0 * (y1 - y2) - x1 * (y0 - y2) + x2 * (y0 - y1)            # det [[x0 x1 x2], [y0 y1 y2], [1 1 1]]
    ok_det = np.abs(det) >= 1e-300
    sdet = np.where(ok_det, det, 1.0)
    idx = -np.ones((size, size), np.int64)
    depth = np.full((size, size), FAR)
    xp = ((2 * np.arange(size) + 1 - size) / size)[:, None]            # [S, 1] against faces [1, F]
    for yi in range(size):
        yp = (2 * yi + 1 - size) / size
        out = ((yp - y0) * (x1 - x0) < (xp - x0) * (y1 - y0)) | ((yp - y1) * (x2 - x1) < (xp - x1) * (y2 - y1)) | \
              ((yp - y2) * (x0 - x2) < (xp - x2) * (y0 - y2))
        # inverse of the 3x3 matrix applied to (xp, yp, 1): the barycentric coordinates
        w0 = ((y1 - y2) * xp + (x2 - x1) * yp + (x1 * y2 - x2 * y1)) / sdet
        w1 = ((y2 - y0) * xp + (x0 - x2) * yp + (x2 * y0 - x0 * y2)) / sdet
        w2 = ((y0 - y1) * xp + (x1 - x0) * yp + (x0 * y1 - x1 * y0)) / sdet
        w0, w1, w2 = np.clip(w0, 0, 1), np.clip(w1, 0, 1), np.clip(w2, 0, 1)
        s = np.maximum(w0 + w1 + w2, 1e-10)
        with np.errstate(divide="ignore", invalid="ignore"):
            zp = 1.0 / ((w0 / z0 + w1 / z1 + w2 / z2) / s)
        valid = (~out) & ok_det[None, :] & (zp > NEAR) & (zp < FAR)
        zp = np.where(valid, zp, FAR)
        best = np.argmin(zp, axis=1)                                    # first minimum = lowest face index on ties
        dmin = zp[np.arange(size), best]
        hit = dmin < FAR
        idx[yi] = np.where(hit, best, -1)
        depth[yi] = np.where(hit, dmin, FAR)
    alpha = (idx >= 0).astype(np.float64)
    return idx, alpha[::-1].copy(), depth[::-1].copy()


def backward_faces(fv, idx, alpha_img, g_alpha_img, size):
    """NMR pseudo-gradient w.r.t. the NDC (x, y) of each face vertex: [2F, 3, 3] (z column stays zero)."""
    alpha, g_alpha = alpha_img[::-1], g_alpha_img[::-1]            # back to the y-up internal layout
    out = np.zeros_like(fv)
    for fn, face in enumerate(fv):
        if (face[2, 1] - face[0, 1]) * (face[1, 0] - face[0, 0]) < (face[1, 1] - face[0, 1]) * (face[2, 0] - face[0, 0]):
            continue
        for edge in range(3):
            pi = [(edge + n) % 3 for n in range(3)]
            pp = np.array([[0.5 * (face[pi[n], d] * size + size - 1) for d in range(2)] for n in range(3)])
            for axis in range(2):
                p = np.array([[pp[n][(d + axis) % 2] for d in range(2)] for n in range(3)])
                if axis == 0:
                    direction = -1 if p[0][0] < p[1][0] else 1
                else:
                    direction = 1 if p[0][0] < p[1][0] else -1
                d0_from = int(max(math.ceil(min(p[0][0], p[1][0])), 0))
                d0_to = int(min(max(p[0][0], p[1][0]), size - 1))
                at = (lambda a, d1, d0: a[d1, d0]) if axis == 0 else (lambda a, d1, d0: a[d0, d1])
                for d0 in range(d0_from, d0_to + 1):
                    d1_cross = (p[1][1] - p[0][1]) / (p[1][0] - p[0][0]) * (d0 - p[0][0]) + p[0][1]
                    d1_in = math.floor(d1_cross) if direction > 0 else math.ceil(d1_cross)
                    d1_out = d1_in + direction
                    if not (0 <= d1_in < size) or not (0 <= d1_out < size):
                        continue
                    alpha_in, alpha_out = at(alpha, d1_in, d0), at(alpha, d1_out, d0)

                    def push(d1, diff_grad):
                        if diff_grad <= 0:
                            return
                        if p[1][0] != d0:
                            dist = (p[1][0] - p[0][0]) / (p[1][0] - d0) * (d1 - d1_cross) * 2.0 / size
                            dist = dist + EPS if dist > 0 else dist - EPS
                            out[fn, pi[0], 1 - axis] -= diff_grad / dist
                        if p[0][0] != d0:
                            dist = (p[1][0] - p[0][0]) / (d0 - p[0][0]) * (d1 - d1_cross) * 2.0 / size
                            dist = dist + EPS if dist > 0 else dist - EPS
                            out[fn, pi[1], 1 - axis] -= diff_grad / dist

                    if at(idx, d1_in, d0) == fn:
                        lim = size - 1 if direction > 0 else 0
                        for d1 in range(max(min(d1_out, lim), 0), min(max(d1_out, lim), size - 1) + 1):
                            push(d1, (at(alpha, d1, d0) - alpha_in) * at(g_alpha, d1, d0))
                    if (d0 - p[0][0]) * (d0 - p[2][0]) < 0:
                        c2 = (p[2][1] - p[0][1]) / (p[2][0] - p[0][0]) * (d0 - p[0][0]) + p[0][1]
                    else:
                        c2 = (p[1][1] - p[2][1]) / (p[1][0] - p[2][0]) * (d0 - p[2][0]) + p[2][1]
                    lim = math.ceil(c2) if direction > 0 else math.floor(c2)
                    for d1 in range(max(min(d1_in, lim), 0), min(max(d1_in, lim), size - 1) + 1):
                        if at(idx, d1, d0) != fn:
                            continue
                        push(d1, (at(alpha, d1, d0) - alpha_out) * at(g_alpha, d1, d0))
    return out


def backward_verts(g_faces, verts, faces, K4):
    """Scatter the face gradients to vertices and chain through ``project``: d loss / d verts [V, 3]."""
    f2 = np.concatenate([faces, faces[:, ::-1]], 0)
    g_ndc = np.zeros((verts.shape[0], 2))
    for fn in range(f2.shape[0]):
        for k in range(3):
            g_ndc[f2[fn, k]] += g_faces[fn, k, :2]
    z = verts[:, 2] + 1e-9
    g = np.zeros_like(verts)
    g[:, 0] = g_ndc[:, 0] * 2 * K4[0] / z
    g[:, 1] = -g_ndc[:, 1] * 2 * K4[1] / z
    g[:, 2] = (-g_ndc[:, 0] * 2 * K4[0] * verts[:, 0] + g_ndc[:, 1] * 2 * K4[1] * verts[:, 1]) / (z * z)
    return g
