"""On-disk outputs of the hot-path stages, in the reference's formats (SURVEY.md section 8(b) "on-disk outputs that later steps read"), so
that the reference's downstream scripts (pack_smplt.py, pack_recon.py, SmoothNet / HVOP-Net loaders, evaluation) read what this package
writes.  Pure host code: one batched device->host copy per call, then the same per-frame files the reference writes.

  save_smplt_fits     ``k{kid}.smplfit_temporal.pkl`` / ``.smplfit_smoothed.pkl`` = {'pose' (156,), 'betas' (10,), 'trans' (3,)}
                      (preprocess/fit_SMPLH_kpts.py:250-261)
  save_smpl_params    ``k{tid}.smpl.pkl`` = {'pose', 'betas', 'trans', 'score'}            (recon/opt_utils.py:134-141)
  save_object_params  ``k{tid}.object.pkl`` = {'rot' (3,3) re-projected to SO(3) without noise, 'trans' (3,), 'scale' ()}
                      (recon/recon_fit_base.py:296-313)
  save_neural_recon   ``k{tid}_densepc.npz`` = {'human': {points, pca_axis, parts, centers, visibility}, 'object': {...}}
                      (recon/recon_fit_base.py:830-844, recon/gen/generator_vis.py:54-55)
  output_folders      ``<outpath>/<seq>/<frame>/<save_name>``                              (recon/recon_fit_base.py:278-294)
  save_triplane_png / load_triplane_png   ``k{kid}.smooth_triplane.png`` with R, G, B = right, back, top (render/render_triplane_nr.py:84-85)
  save_ply / load_ply ``k{kid}.smplfit_*.ply`` meshes (psbody ``Mesh.write_ply``; read back by the triplane renderer and the test loader)
  save_infill_output  the pack HVOP-Net hands to the joint optimisation (interp/test_infiller.py:129-143)
  packed_batch        the slices of a sequence pack for the frames of one batch (recon/recon_fit_base.py:346-370)
  pack_smplt / pack_recon / load_packed   the per-sequence joblib packs (preprocess/pack_smplt.py:45-63, preprocess/pack_recon.py:118-157)
                      written from in-memory trajectories instead of re-reading every per-frame file
"""
from __future__ import annotations

import os
import pickle as pkl
from typing import Dict, List, Sequence

import numpy as np
import torch


def _np(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def output_folders(outpath: str, image_paths: Sequence[str], save_name: str) -> List[str]:
    """ROOT/SEQ/frame_time/kx.color.jpg -> <outpath>/SEQ/frame_time/<save_name> (created)."""
    folders = []
    for x in image_paths:
        parts = str(x).split(os.sep)
        folder = os.path.join(outpath, parts[-3], parts[-2], save_name)
        os.makedirs(folder, exist_ok=True)
        folders.append(folder)
    return folders


def save_smplt_fits(outfiles: Sequence[str], poses, betas, trans, skip=None) -> int:
    """One ``{'pose', 'betas', 'trans'}`` pickle per frame (the SMPL-T pre-fit / smoothed-fit result); frames flagged in `skip` are not
    written, as ``BaseFitter.save_results`` skips frames without confident keypoints.  Returns the number of files written."""
    poses, betas, trans = _np(poses), _np(betas), _np(trans)
    n = 0
    for i, f in enumerate(outfiles):
        if skip is not None and bool(skip[i]):
            continue
        os.makedirs(os.path.dirname(f) or ".", exist_ok=True)
        with open(f, "wb") as fh:
            pkl.dump({"pose": poses[i], "betas": betas[i], "trans": trans[i]}, fh)
        n += 1
    return n


def save_smpl_params(folders: Sequence[str], tid: int, pose, betas, trans, scores=None) -> List[str]:
    poses, betas, trans = _np(pose), _np(betas), _np(trans)
    scores = np.zeros(len(folders)) if scores is None else _np(scores)
    files = []
    for i, folder in enumerate(folders):
        f = os.path.join(folder, f"k{tid}.smpl.pkl")
        with open(f, "wb") as fh:
            pkl.dump({"pose": poses[i], "betas": betas[i], "trans": trans[i], "score": scores[i]}, fh)
        files.append(f)
    return files


def save_object_params(folders: Sequence[str], tid: int, obj_R, obj_t, obj_s) -> List[str]:
    """`obj_R` is the free 3x3 optimisation variable: it is projected to SO(3) WITHOUT the decopose_axis noise before saving
    (recon/recon_fit_base.py:303, `no_rand=True`) -- always, as the reference does: a host array / CPU tensor goes through the same SVD
    projection on the host (U diag(1, 1, det(U V^T)) V^T, what ``decopose_axis`` computes)."""
    from .geom import project_so3
    if torch.is_tensor(obj_R) and obj_R.is_cuda:
        R = _np(project_so3(obj_R.detach()))
    else:
        A = np.asarray(_np(obj_R), np.float64).reshape(-1, 3, 3)
        U, _, Vt = np.linalg.svd(A)
        d = np.linalg.det(U @ Vt)
        D = np.tile(np.eye(3), (A.shape[0], 1, 1)); D[:, 2, 2] = d
        R = (U @ D @ Vt).astype(np.float32)
    t, s = _np(obj_t), _np(obj_s)
    files = []
    for i, folder in enumerate(folders):
        f = os.path.join(folder, f"k{tid}.object.pkl")
        with open(f, "wb") as fh:
            pkl.dump({"rot": R[i], "trans": t[i], "scale": s[i]}, fh)
        files.append(f)
    return files


def save_neural_recon(folders: Sequence[str], tid: int, recon_batch: Dict[str, Dict[str, torch.Tensor]]) -> List[str]:
    """recon_batch = {'human': {'points' [B, n, 3], 'pca_axis', 'parts', 'centers', 'visibility'}, 'object': {...}} as
    ``GeneratorTriplaneVis.generate_pclouds_batch`` returns it; one npz per frame holding the two per-target dicts."""
    host = {tar: {k: _np(v) for k, v in d.items()} for tar, d in recon_batch.items()}
    files = []
    for i, folder in enumerate(folders):
        f = os.path.join(folder, f"k{tid}_densepc.npz")
        np.savez(f, **{tar: {k: v[i] for k, v in d.items()} for tar, d in host.items()})
        files.append(f)
    return files


def load_smplt_fits(files: Sequence[str]):
    """Inverse of save_smplt_fits: stacked (poses [T, 156], betas [T, 10], trans [T, 3]) -- what SMPLTSmoother.load_inputs_raw collects
    frame by frame (smoothnet/smooth_smplt.py:130-152)."""
    P, B, T = [], [], []
    for f in files:
        with open(f, "rb") as fh:
            d = pkl.load(fh)
        P.append(d["pose"]); B.append(d["betas"]); T.append(d["trans"])
    return np.stack(P, 0), np.stack(B, 0), np.stack(T, 0)


def pack_smplt(outfile: str, frames: Sequence[str], gender: str, poses, betas, trans) -> str:
    """The per-sequence pack ``preprocess/pack_smplt.py:45-63`` writes from the per-frame SMPL-T fits -- here straight from the (all-gathered)
    trajectory: {'poses' (T,156), 'betas', 'trans', dummy 'obj_angles' eye / 'obj_trans' zeros / 'obj_scales' zeros, 'gender', 'frames'}."""
    import joblib
    poses, betas, trans = _np(poses), _np(betas), _np(trans)
    L = poses.shape[0]
    os.makedirs(os.path.dirname(outfile) or ".", exist_ok=True)
    joblib.dump({"poses": poses, "betas": betas, "trans": trans, "obj_angles": np.eye(3)[None].repeat(L, 0), "obj_trans": np.zeros((L, 3)),
                 "obj_scales": np.zeros((L,)), "gender": gender, "frames": list(frames)}, outfile)
    return outfile


def pack_recon(outfile: str, frames: Sequence[str], gender: str, recon_name: str, neural_pca, neural_trans, neural_visibility, poses=None, betas=None,
               trans=None, root_joints=None, obj_angles=None, obj_trans=None, obj_scales=None) -> str:
    """The per-sequence pack of ``preprocess/pack_recon.py:29-161`` from in-memory trajectories.  Without the SMPL / object arrays it is the
    ``-neural_only`` pack (step 4.1 of demo.sh: 'neural_pca' / 'neural_trans' / 'neural_visibility' lists, 'recon_exist', meta); with them
    the full one ('poses' Tx156|72, 'betas', 'trans', 'root_joints', 'obj_angles' as stored = R^T, 'obj_trans', 'obj_scales').
    ``neural_trans`` is the object centre relative to the body (``centers[3:]``)."""
    import joblib
    T = len(frames)
    d = {}
    full = poses is not None
    if full:
        d.update({"poses": _np(poses), "betas": _np(betas), "trans": _np(trans), "root_joints": _np(root_joints), "obj_angles": _np(obj_angles),
                  "obj_trans": _np(obj_trans), "obj_scales": np.asarray(_np(obj_scales)).reshape(T)})
    d.update({"neural_pca": [a for a in _np(neural_pca).reshape(T, 3, 3)], "neural_trans": [a for a in _np(neural_trans).reshape(T, 3)],
              "neural_visibility": [a for a in _np(neural_visibility).reshape(T, -1)], "recon_exist": np.ones(T, bool), "recon_name": recon_name,
              "frames": list(frames), "gender": gender})
    if not full:                                                                   # key order of the reference's neural-only dict
        d = {k: d[k] for k in ("neural_pca", "neural_trans", "recon_exist", "neural_visibility", "recon_name", "frames", "gender")}
    os.makedirs(os.path.dirname(outfile) or ".", exist_ok=True)
    joblib.dump(d, outfile)
    return outfile


def load_packed(file: str) -> dict:
    """``joblib.load`` of a pack written by the reference or by pack_smplt / pack_recon, list entries stacked: what ``ReconDataReader`` users,
    ``ObjrotSmoother.load_inputs_raw`` (smooth_objrot.py:36-69) and ``MotionInfillAutoreg.test`` (test_infill_autoreg.py:60-82) start from."""
    import joblib
    d = dict(joblib.load(file))
    for k in ("neural_pca", "neural_trans", "neural_visibility"):
        if k in d and isinstance(d[k], list) and len(d[k]):
            d[k] = np.stack([np.asarray(a) for a in d[k]], 0)
    return d


def save_triplane_png(files: Sequence[str], masks) -> List[str]:
    """``k{kid}.smooth_triplane.png`` and friends (render/render_triplane_nr.py:84-85): masks [B, 3, S, S] (right, back, top; 0/1 or bool, as
    ``TriplaneNrRenderer.render_3views`` returns them) -> 8-bit RGB png with R = right, G = back, B = top scaled to 0/255 -- the bytes
    ``cv2.imwrite(outfile, stack[:, :, ::-1])`` puts on disk."""
    from PIL import Image
    m = _np(masks)
    if m.ndim != 4 or m.shape[1] != 3 or len(files) != m.shape[0]:
        raise ValueError(f"expected masks [B, 3, S, S] and B file names, got {m.shape} and {len(files)}")
    rgb = ((m > 0).astype(np.uint8) * 255).transpose(0, 2, 3, 1)
    for f, img in zip(files, rgb):
        os.makedirs(os.path.dirname(f) or ".", exist_ok=True)
        Image.fromarray(np.ascontiguousarray(img), "RGB").save(f)
    return list(files)


def load_triplane_png(file: str) -> np.ndarray:
    """[S, S, 3] uint8 in (right, back, top) order: ``cv2.imread(file)[:, :, ::-1]`` of data/testdata_triplane.py:79 (before the / 255)."""
    from PIL import Image
    return np.asarray(Image.open(file).convert("RGB"))


def packed_batch(packed: dict, image_paths: Sequence[str], test_kid: int = 1) -> Dict[str, np.ndarray]:
    """The slices of a sequence pack that belong to the frames of one batch -- ``extract_frame_inds`` + ``load_old_smpl_recon`` /
    ``load_old_obj_recon`` / ``load_occ_ratios_recon`` (recon/recon_fit_base.py:346-370, recon/recon_fit_triplane.py:146-174): image paths
    ``<root>/<seq>/<frame time>/k<kid>.color.jpg`` are matched against ``packed['frames']``; every file must come from ``test_kid``.
    Returns the per-frame entries present in the pack ('poses', 'betas', 'trans', 'obj_angles', 'obj_trans', 'obj_scales', 'neural_pca',
    'neural_trans', 'occ_ratios' = neural_visibility[:, 0]) stacked over the batch, plus 'frame_inds'."""
    frames = list(packed["frames"])
    inds = []
    for f in image_paths:
        name = os.path.basename(f)
        kid = int(name.split(".")[0][1])
        assert kid == test_kid, f"{f} kinect id={kid}!= test kid={test_kid}"
        inds.append(frames.index(os.path.basename(os.path.dirname(f))))
    out = {"frame_inds": np.asarray(inds)}
    for k in ("poses", "betas", "trans", "obj_angles", "obj_trans", "obj_scales", "neural_pca", "neural_trans"):
        if k in packed and len(packed[k]):
            out[k] = np.stack([np.asarray(packed[k][i]) for i in inds], 0)
    if "neural_visibility" in packed and len(packed["neural_visibility"]):
        out["occ_ratios"] = np.array([np.asarray(packed["neural_visibility"][i]).reshape(-1)[0] for i in inds])
    return out


def save_ply(file: str, verts, faces) -> str:
    """``Mesh(v, f).write_ply(file)`` (psbody.mesh; called at preprocess/fit_SMPLH_kpts.py:262-264, fit_SMPLH_smoothed.py:62-63): the
    ``k{kid}.smplfit_temporal.ply`` / ``.smplfit_smoothed.ply`` meshes that the triplane renderer and ``TestDataTriplane.load_mesh`` read back.
    Binary little-endian PLY, float32 x y z per vertex, ``list uchar int vertex_indices`` per face."""
    v = np.ascontiguousarray(_np(verts), dtype="<f4").reshape(-1, 3)
    f = np.ascontiguousarray(_np(faces), dtype="<i4").reshape(-1, 3)
    os.makedirs(os.path.dirname(file) or ".", exist_ok=True)
    header = ("ply\nformat binary_little_endian 1.0\n"
              f"element vertex {v.shape[0]}\nproperty float x\nproperty float y\nproperty float z\n"
              f"element face {f.shape[0]}\nproperty list uchar int vertex_indices\nend_header\n")
    rec = np.empty(f.shape[0], dtype=[("n", "u1"), ("idx", "<i4", (3,))])
    rec["n"], rec["idx"] = 3, f
    with open(file, "wb") as fh:
        fh.write(header.encode("ascii")); fh.write(v.tobytes()); fh.write(rec.tobytes())
    return file


def load_ply(file: str):
    """(verts [V,3] float64, faces [F,3] int64) of an ascii or binary-little-endian triangle PLY whose vertex element starts with float x, y, z
    (further float / uchar vertex properties are skipped) -- enough for the meshes the reference's fitters write."""
    sizes = {"float": 4, "float32": 4, "double": 8, "float64": 8, "uchar": 1, "uint8": 1, "char": 1, "int": 4, "int32": 4, "uint": 4, "short": 2, "ushort": 2}
    with open(file, "rb") as fh:
        fmt, nv, nf, vprops, cur = None, 0, 0, [], None
        while True:
            line = fh.readline().decode("ascii").strip()
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                cur = tok[1]
                if cur == "vertex":
                    nv = int(tok[2])
                elif cur == "face":
                    nf = int(tok[2])
            elif tok[0] == "property" and cur == "vertex":
                vprops.append((tok[2], tok[1]))
            elif tok[0] == "end_header":
                break
        if [p[0] for p in vprops[:3]] != ["x", "y", "z"]:
            raise ValueError(f"{file}: vertex element does not start with x, y, z")
        if fmt == "ascii":
            rows = [fh.readline().split() for _ in range(nv)]
            verts = np.array([[float(t) for t in r[:3]] for r in rows], np.float64)
            faces = np.array([[int(t) for t in fh.readline().split()[1:4]] for _ in range(nf)], np.int64)
        elif fmt == "binary_little_endian":
            vdt = np.dtype([(n, {4: "<f4", 8: "<f8"}[sizes[t]] if t in ("float", "float32", "double", "float64") else f"<u{sizes[t]}") for n, t in vprops])
            vraw = np.frombuffer(fh.read(vdt.itemsize * nv), dtype=vdt, count=nv)
            verts = np.stack([vraw["x"], vraw["y"], vraw["z"]], 1).astype(np.float64)
            fdt = np.dtype([("n", "u1"), ("idx", "<i4", (3,))])
            fraw = np.frombuffer(fh.read(fdt.itemsize * nf), dtype=fdt, count=nf)
            if nf and not bool((fraw["n"] == 3).all()):
                raise ValueError(f"{file}: only triangle meshes are supported")
            faces = fraw["idx"].astype(np.int64)
        else:
            raise ValueError(f"{file}: unsupported PLY format {fmt}")
    return verts, faces.reshape(-1, 3)


def save_infill_output(dat: dict, outfile: str, obj_angles=None, obj_trans=None, exp_name: str = "cmf-k4-lrot") -> str:
    """``MotionInfillTester.save_output`` (interp/test_infiller.py:129-143): the pack the HVOP-Net stage hands to the joint optimisation
    (``-or smooth-hvopnet``).  ``dat`` is the input pack (``load_packed``); ``obj_angles`` [T,3,3] as the in-filler returns them (= R^T, what
    the reference stores after its transpose) and ``obj_trans`` replace the pack's entries -- ``None`` keeps the input (the reference's
    ``save_orig`` branch when the first clip has no visible seed frames); 'obj_scales' becomes ones and 'exp_name' is added."""
    import joblib
    d = dict(dat)
    L = len(d["frames"])
    if obj_angles is not None:
        d["obj_angles"] = _np(obj_angles).copy()
        if obj_trans is not None:
            d["obj_trans"] = _np(obj_trans)
    d["obj_scales"] = np.ones(L)
    d["exp_name"] = exp_name
    os.makedirs(os.path.dirname(outfile) or ".", exist_ok=True)
    joblib.dump(d, outfile)
    return outfile
