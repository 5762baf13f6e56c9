"""CPU restatement of SIF-Net inference (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Functional (no nn.Module): every routine takes the reference ``state_dict`` and a key prefix, so the same
checkpoint drives the reference classes, this oracle and the CUDA path.  Works in fp32 or fp64 (pass
``dtype=torch.float64`` to separate kernel error from the reference's own fp32 rounding).

Follows: model/HGFilters.py:26-50,162-203 · model/net_util.py:374-396 · model/chore.py:113-144 ·
model/chore_triplane.py:60-164,207-251 · model/chore_tri_vis.py:31-50 · model/geometry.py:4-14 ·
model/camera.py:45-89.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _gn_relu(sd: SD, key: str, x: torch.Tensor) -> torch.Tensor:
    # nn.GroupNorm(32, C) (eps 1e-5, biased variance) followed by ReLU -- net_util.py:358-362,377-386
    return F.relu(F.group_norm(x, 32, sd[key + ".weight"].to(x.dtype), sd[key + ".bias"].to(x.dtype), 1e-5))


def conv_block(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """ConvBlock.forward, model/net_util.py:374-396: three GN->ReLU->3x3 convs whose outputs are
    concatenated (1/2, 1/4, 1/4 of the channels) and added to the (optionally 1x1-projected) input."""
    w = lambda k: sd[f"{p}.{k}.weight"].to(x.dtype)
    o1 = F.conv2d(_gn_relu(sd, f"{p}.bn1", x), w("conv1"), padding=1)
    o2 = F.conv2d(_gn_relu(sd, f"{p}.bn2", o1), w("conv2"), padding=1)
    o3 = F.conv2d(_gn_relu(sd, f"{p}.bn3", o2), w("conv3"), padding=1)
    if f"{p}.downsample.2.weight" in sd:       # in_planes != out_planes, net_util.py:364-370
        res = F.conv2d(_gn_relu(sd, f"{p}.bn4", x), sd[f"{p}.downsample.2.weight"].to(x.dtype))
    else:
        res = x
    return torch.cat((o1, o2, o3), 1) + res


def hourglass(sd: SD, p: str, level: int, x: torch.Tensor) -> torch.Tensor:
    """HourGlass._forward, model/HGFilters.py:26-50."""
    up1 = conv_block(sd, f"{p}.b1_{level}", x)
    low = conv_block(sd, f"{p}.b2_{level}", F.avg_pool2d(x, 2, stride=2))
    if level > 1:
        low = hourglass(sd, p, level - 1, low)
    else:
        low = conv_block(sd, f"{p}.b2_plus_{level}", low)
    low = conv_block(sd, f"{p}.b3_{level}", low)
    up2 = F.interpolate(low, scale_factor=2, mode="bicubic", align_corners=True)
    return up1 + up2


def hg_filter(sd: SD, p: str, x: torch.Tensor, num_stack: int, depth: int) -> Tuple[List[torch.Tensor], torch.Tensor, torch.Tensor]:
    """HGFilter.forward, model/HGFilters.py:162-203.  Returns (outputs per stack, tmpx, normx)."""
    c = lambda k, t: F.conv2d(t, sd[f"{p}.{k}.weight"].to(t.dtype), sd[f"{p}.{k}.bias"].to(t.dtype))
    x = _gn_relu(sd, f"{p}.bn1", F.conv2d(x, sd[f"{p}.conv1.weight"].to(x.dtype), sd[f"{p}.conv1.bias"].to(x.dtype),
                                          stride=2, padding=3))
    tmpx = x
    x = F.avg_pool2d(conv_block(sd, f"{p}.conv2", x), 2, stride=2)
    normx = x
    x = conv_block(sd, f"{p}.conv4", conv_block(sd, f"{p}.conv3", x))
    previous, outputs = x, []
    for i in range(num_stack):
        ll = conv_block(sd, f"{p}.top_m_{i}", hourglass(sd, f"{p}.m{i}", depth, previous))
        ll = _gn_relu(sd, f"{p}.bn_end{i}", c(f"conv_last{i}", ll))
        out = c(f"l{i}", ll)
        outputs.append(out)
        if i < num_stack - 1:
            previous = previous + c(f"bl{i}", ll) + c(f"al{i}", out)
    return outputs, tmpx, normx


def sif_filter(sd: SD, images: torch.Tensor, num_stack: int = 3, depth: int = 2) -> Dict[str, object]:
    """CHORETriplane.filter in eval mode (model/chore_triplane.py:60-95, model/chore.py:128-144): RGB+masks
    encoder on channels 0-4, the shared triplane encoder on channels 5, 6, 7; only the last stack is kept."""
    assert images.shape[1] == 8, f"given image shape invalide: {images.shape}"
    outs, tmpx, _ = hg_filter(sd, "image_filter", images[:, :5], num_stack, depth)
    tri_feat, tri_tmpx = [], []
    for v in range(3):
        o, t, _ = hg_filter(sd, "triplane_encoder", images[:, 5 + v:6 + v], num_stack, depth)
        tri_feat.append(o[-1]); tri_tmpx.append(t)
    return {"im_feat": outs[-1], "tmpx": tmpx, "tri_feat": tri_feat, "tri_tmpx": tri_tmpx}


def project_points(points: torch.Tensor, crop_center: torch.Tensor, fx: float, fy: float, cx: float, cy: float,
                   crop: float) -> torch.Tensor:
    """KinectColorCamera.project_points with an offset, model/camera.py:45-82: pinhole projection into the
    2048x1536 image, re-centred on the crop and normalised to [-1, 1].  Returns [B, 2, N]."""
    x, y, z = points[..., 0], points[..., 1], points[..., 2]
    px = fx * x / z + cx
    py = fy * y / z + cy
    px = crop / 2 + px - crop_center[:, 0:1]
    py = crop / 2 + py - crop_center[:, 1:2]
    return torch.stack([2 * px / crop - 1, 2 * py / crop - 1], 1)


def triplane_uv(points: torch.Tensor, body_center: torch.Tensor) -> List[torch.Tensor]:
    """CHORETriplane.triplane_project, model/chore_triplane.py:220-251 (fx=1, cx=0): right (z, y),
    back (-x, y), top (x, -z) of the body-centred point.  Each [B, 2, N]."""
    c = points - body_center[:, None, :]
    return [torch.stack([c[..., 2], c[..., 1]], 1), torch.stack([-c[..., 0], c[..., 1]], 1),
            torch.stack([c[..., 0], -c[..., 2]], 1)]


def sample(feat: torch.Tensor, uv: torch.Tensor) -> torch.Tensor:
    """model/geometry.py:4-14: bilinear grid_sample, zeros padding, align_corners=True.  [B,C,N]."""
    return F.grid_sample(feat, uv.transpose(1, 2).unsqueeze(2), align_corners=True)[..., 0]


def point_features(maps: Dict[str, object], points, crop_center, body_center, cam) -> Tuple[torch.Tensor, torch.Tensor]:
    """CHORETriplane.query_features, model/chore_triplane.py:166-205.  Channel order:
    im_feat | x, y, z-2.2 | tmpx | tri_tmpx right, back, top | tri_feat right | back | top."""
    xy = project_points(points, crop_center, *cam)
    zf = torch.stack([points[..., 0], points[..., 1], points[..., 2] - 2.2], 1)   # get_zfeat, :207-218
    uvs = triplane_uv(points, body_center)
    parts = [sample(maps["im_feat"], xy), zf, sample(maps["tmpx"], xy)]
    parts += [sample(t, uv) for t, uv in zip(maps["tri_tmpx"], uvs)]
    parts += [sample(f, uv) for f, uv in zip(maps["tri_feat"], uvs)]
    return torch.cat(parts, 1), xy


def mlp_head(sd: SD, name: str, feat: torch.Tensor) -> torch.Tensor:
    """make_decoder, model/chore.py:113-126: Conv1d(k=1) F->H->H->H->out with ReLU between."""
    h = feat
    for idx in (0, 2, 4, 6):
        h = F.conv1d(h, sd[f"{name}.{idx}.weight"].to(h.dtype), sd[f"{name}.{idx}.bias"].to(h.dtype))
        if idx != 6:
            h = F.relu(h)
    return h


def sif_query(sd: SD, maps, points, crop_center, body_center, cam, out_dist: float = 5.0):
    """CHORETriplane.query + CHORETriplaneVisibility.decode (model/chore_triplane.py:97-164,
    model/chore_tri_vis.py:31-50).  Returns (df[B,2,N], pca[B,3,3,N], parts[B,14,N], centers[B,3,N], vis[B,1,N])."""
    feat, xy = point_features(maps, points, crop_center, body_center, cam)
    df = mlp_head(sd, "df", feat)
    pca = mlp_head(sd, "pca_predictor", feat)
    parts = mlp_head(sd, "part_predictor", feat)
    centers = mlp_head(sd, "center_predictor", feat)
    vis = torch.sigmoid(mlp_head(sd, "visib_predictor", feat))
    in_img = (xy[:, 0] >= -1.0) & (xy[:, 0] <= 1.0) & (xy[:, 1] >= -1.0) & (xy[:, 1] <= 1.0)
    df = torch.where(in_img[:, None, :], df, torch.full_like(df, out_dist))      # chore_triplane.py:156-159
    return df, pca.view(df.shape[0], 3, 3, -1), parts, centers, vis
