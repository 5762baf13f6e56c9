"""The per-batch body of ``ReconFitterTriplane.fit_recon`` (recon/recon_fit_triplane.py:47-111) on in-memory tensors: neural reconstruction in
mini-batches, filter of the whole batch, SMPL refinement, object initialisation, joint optimisation.  Everything file-shaped (data loader,
``get_smpl_init`` / ``load_old_obj_recon`` pickles, key-point json, ``save_outputs``) is an argument or a return value, so the reference's
``fit_recon`` loop -- or a rank of ``parallel.rank_frames`` -- calls one function per 96-frame batch and writes the results with
``vistracker_b200.io``.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

from ._lib import nvtx_range
from .generator import GeneratorTriplaneVis
from .geom import init_object_orientation
from .recon_fit import SMPL_POSE_PRAMS_NUM, ReconFitterTriVisFull, SMPLParams

MINI_BATCH = 16            # recon_fit_behave.py:124 (8 on the 'volta' partition)


def combine_mini_batches(pcs, samples_count: int) -> Dict[str, Dict[str, torch.Tensor]]:
    """recon_fit_behave.py:152-183: points / parts truncated to the smallest per-mini-batch sample count, everything else concatenated."""
    out = {}
    for target in ("human", "object"):
        out[target] = {}
        for key in pcs[0][target]:
            parts = [(pc[target][key][:, :samples_count] if key in ("points", "parts") else pc[target][key]) for pc in pcs]
            out[target][key] = torch.cat(parts, 0)
    return out


def generate_all(generator: GeneratorTriplaneVis, data: Dict[str, torch.Tensor], num_points: int = 4000, mini_batch_size: int = MINI_BATCH,
                 on_mini_batch: Optional[Callable] = None, keep_maps: bool = False):
    """``ReconFitterBehave.generate_all`` (recon_fit_behave.py:121-150): ``generate_pclouds_batch`` per mini-batch of 16 frames with 10
    projection steps, then ``combine_mini_batches``.  ``on_mini_batch(start, end, pc_generated)`` is where the reference writes
    ``k{tid}_densepc.npz`` (``save_neural_recon``).

    ``keep_maps``: every mini-batch's feature maps are copied into whole-batch tensors as they are produced and left in ``net._maps`` at
    the end, i.e. exactly what ``filter_batch`` would recompute.  The reference filters the whole batch a second time after the generator
    (recon_fit_triplane.py:57-60: the generator's per-mini-batch filter calls overwrite the network's cached maps and the whole batch does
    not fit its GPU otherwise); the maps are a pure function of the frames, so keeping them gives the same values for half the encoder work."""
    B = data["images"].shape[0]
    net = generator.model
    pcs, samples, full = [], 100000, None
    for s in range(0, B, mini_batch_size):
        mini = {k: v[s:s + mini_batch_size] for k, v in data.items()}
        pc = generator.generate_pclouds_batch(mini, num_points=num_points, num_steps=10, mute=True)
        if keep_maps and B > mini_batch_size:
            m, n = net._maps, mini["images"].shape[0]
            if full is None:
                full = tuple(torch.empty((B if i < 2 else 3 * B,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for i, t in enumerate(m))
            full[0][s:s + n].copy_(m[0]); full[1][s:s + n].copy_(m[1])
            for v in range(3):
                full[2][v * B + s:v * B + s + n].copy_(m[2][v * n:(v + 1) * n]); full[3][v * B + s:v * B + s + n].copy_(m[3][v * n:(v + 1) * n])
        if on_mini_batch is not None:
            on_mini_batch(s, min(s + mini_batch_size, B), pc)
        samples = min(samples, pc["human"]["points"].shape[1], pc["object"]["points"].shape[1])
        pcs.append(pc)
    if full is not None:
        net._maps = full
    return combine_mini_batches(pcs, samples)


def filter_batch(net, images: torch.Tensor, chunk: int = MINI_BATCH) -> None:
    """``generator.filter(data)`` on the whole optimisation batch (recon_fit_triplane.py:58-60).  The encoders run in chunks to bound the
    activation memory; the cached maps end up covering all B frames (image / tmpx maps concatenated along the batch, the three triplane
    views kept view-major as ``filter`` lays them out)."""
    B = images.shape[0]
    if B <= chunk:
        net.filter(images.to(net.device))
        return
    direct = net.use_graph and net.filter_streams == 1          # graph replay: each chunk's maps are copied straight into their rows
    full, parts = None, []
    for s in range(0, B, chunk):
        part = images[s:s + chunk].to(net.device)
        n = part.shape[0]
        if direct and full is not None:
            net.filter(part, into=(full, s, B))
            continue
        net.filter(part)
        if not direct:
            parts.append((net._maps, n))
            continue
        m = net._maps                                            # first chunk: learn the map shapes, allocate the whole-batch tensors once
        full = tuple(torch.empty((B if i < 2 else 3 * B,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for i, t in enumerate(m))
        full[0][:n].copy_(m[0]); full[1][:n].copy_(m[1])
        for v in range(3):
            full[2][v * B:v * B + n].copy_(m[2][v * n:(v + 1) * n]); full[3][v * B:v * B + n].copy_(m[3][v * n:(v + 1) * n])
    if not direct:                                               # eager filter (VT_FILTER_GRAPH=0 / multi-stream experiments): concatenate
        cat_view = lambda i: torch.cat([torch.cat([m[i][v * n:(v + 1) * n] for m, n in parts]) for v in range(3)])
        full = (torch.cat([m[0] for m, _ in parts]), torch.cat([m[1] for m, _ in parts]), cat_view(2), cat_view(3))
    net._maps = full


def scale_body_kpts(kpts: torch.Tensor, crop_center: torch.Tensor, crop_size: float = 1200.0, net_in_size: float = 512.0,
                    resize_scale: Optional[torch.Tensor] = None, crop_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``scale_body_kpts`` (recon_fit_base.py:397-409): OpenPose key points [B,25,3] from original-image pixels to network-input pixels."""
    B = kpts.shape[0]
    one = torch.ones(B, device=kpts.device)
    rs, cs = (one if resize_scale is None else resize_scale), (one if crop_scale is None else crop_scale)
    pxy = kpts[:, :, :2] * rs[:, None, None]
    size = cs * crop_size
    pxy = pxy - crop_center[:, None, :] + size[:, None, None] / 2
    pxy = pxy * net_in_size / size[:, None, None]
    return torch.cat([pxy, kpts[:, :, 2:3]], -1)


def fit_recon_batch(fitter: ReconFitterTriVisFull, generator: GeneratorTriplaneVis, data: Dict[str, torch.Tensor], smpl_init: Callable,
                    body_kpts: torch.Tensor, obj_points: torch.Tensor, pca_init: Optional[torch.Tensor] = None,
                    obj_rot_init: Optional[torch.Tensor] = None, silhouette=None, occ_ratios: Optional[torch.Tensor] = None,
                    neural_only: bool = False, on_mini_batch: Optional[Callable] = None, noise_fn=None, mini_batch_size: int = MINI_BATCH,
                    reuse_generator_maps: bool = True, **loop_kw):
    """One batch of ``fit_recon``.

    data: {'images' [B,8,512,512], 'crop_center' [B,2], 'body_center' [B,3]} (device or host tensors).
    smpl_init(human_t [B,3]) -> SMPLParams: ``get_smpl_init`` (recon_fit_trivis_full.py:62-77) -- the smoothed SMPL-T parameters of the frames
    (the tri-vis fitter keeps their own translation and ignores ``human_t`` = the pre-fit body centre; it is passed for the variants that do not).
    body_kpts [B,25,3]: OpenPose key points already in network-input pixels (``scale_body_kpts``).
    obj_points [n,3]: surface samples of the object template (``compute_pca_init``); pca_init [3,3]: its PCA axes, used when the
    rotation comes from the network (``-or neural``); obj_rot_init [B,3,3]: the rotation loaded from an earlier stage (HVOP-Net) instead.
    reuse_generator_maps: keep the maps the generator's per-mini-batch filter calls produced instead of filtering the whole batch again (same
    values, see ``generate_all``).
    silhouette: a ``render.SilLossROI`` for the batch, occ_ratios [B]: visibility used by the occlusion-aware terms (defaults to the
    network's prediction, recon_fit_triplane.py:68).
    Returns {'pc_generated', 'smpl', 'obj_R' (projected, no noise), 'obj_t', 'obj_s', 'hist_smpl', 'hist_obj', 'stopped_*', 'smpl_scale'} -- or only 'pc_generated' for
    ``neural_only`` (demo.sh step 4)."""
    dev = fitter.model.device
    B = data["images"].shape[0]
    reuse = reuse_generator_maps and not neural_only and generator.model is fitter.model
    with nvtx_range("fit_recon_batch/generate_all"):
        pc = generate_all(generator, data, on_mini_batch=on_mini_batch, mini_batch_size=mini_batch_size, keep_maps=reuse)
    if neural_only:
        return {"pc_generated": pc}
    if silhouette is None and fitter.scan is None:
        raise ValueError("the 'sil' phase needs a render.SilLossROI for the batch, or fitter.scan = (template vertices, faces) to build one")
    if not reuse:                                    # "need to run image filter again" (recon_fit_triplane.py:57-60)
        with torch.no_grad(), nvtx_range("fit_recon_batch/filter"):
            filter_batch(fitter.model, data["images"], chunk=mini_batch_size)
    human_t = data["body_center"].to(dev).float()                                   # ReconFitterTriplane.get_smpl_translation (recon_fit_triplane.py:210-220):
                                                                                    # the pre-fit body centre, not the network's prediction
    smpl = smpl_init(human_t)
    query_dict = {"crop_center": data["crop_center"].to(dev), "body_center": data["body_center"].to(dev)}
    dd = {"part_labels": fitter.part_labels.to(dev)[None].repeat(B, 1) if fitter.part_labels.dim() == 1 else fitter.part_labels.to(dev),
          "query_dict": query_dict, "body_kpts": body_kpts.float().to(dev),
          "pose_init": smpl.pose[:, 3:SMPL_POSE_PRAMS_NUM].detach().clone().to(dev)}
    with nvtx_range("fit_recon_batch/optimize_smpl"):
        smpl, scale = fitter.optimize_smpl(smpl, dd, iter_for_kpts=1, iter_for_pose=1, iter_for_betas=1, **loop_kw)   # recon_fit_triplane.py:66
    hist_smpl, stopped_smpl = fitter.last_hist, fitter.last_stopped
    # init_obj_fit_data (recon_fit_trivis_full.py:79-104): predicted centre relative to the optimised body centre, rotation from the PCA axes
    obj_t = (pc["object"]["centers"][:, 3:].to(dev) + human_t.to(dev)).detach().clone().requires_grad_(True)
    if obj_rot_init is None:
        if pca_init is None:
            raise ValueError("either obj_rot_init or pca_init (the template's PCA axes) is needed")
        obj_R = init_object_orientation(pc["object"]["pca_axis"].to(dev), pca_init.to(dev), noise=None if noise_fn is None else noise_fn())
    else:
        obj_R = obj_rot_init.to(dev).float()
    obj_R = obj_R.detach().clone().requires_grad_(True)
    obj_s = torch.ones(B, device=dev)
    vis = pc["object"]["visibility"].to(dev).reshape(B) if occ_ratios is None else occ_ratios.to(dev)
    dd.update({"obj_R": obj_R, "obj_t": obj_t, "obj_s": obj_s, "objects": obj_points.to(dev)[None].repeat(B, 1, 1), "occ_ratios": vis,
               "silhouette": silhouette, "trans_init": obj_t.detach().clone(), "smpl": smpl, "images": data["images"]})
    with nvtx_range("fit_recon_batch/optimize_smpl_object"):
        smpl, obj_R, obj_t = fitter.optimize_smpl_object(fitter.model, dd, noise_fn=noise_fn, **loop_kw)                 # recon_fit_triplane.py:106
    return {"pc_generated": pc, "smpl": smpl, "obj_R": fitter.final_rotation(obj_R), "obj_t": obj_t.detach(), "obj_s": obj_s, "hist_smpl": hist_smpl,
            "hist_obj": fitter.last_hist, "stopped_smpl": stopped_smpl, "stopped_obj": fitter.last_stopped, "smpl_scale": scale}
