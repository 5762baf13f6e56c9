"""Where the shims get their models from.  The reference loads them through PATHS.yml: the SMPL-H pickles (SMPL_MODEL_ROOT), its ``assets/``
folder (priors, regressors, part labels), ``experiments/<exp_name>/checkpoints`` for SIF-Net and the BEHAVE object templates.  An
``AssetProvider`` hides that behind five calls; ``ReferenceAssets`` reads the reference's files, ``SyntheticAssets`` builds the seeded
random-init stand-ins this repository tests and benches with (there is no network access for the real ones).
"""
from __future__ import annotations

import glob
import os
import pickle as pkl
from typing import Dict, Tuple

import numpy as np
import torch

from .. import CHORETriplaneVisibility
from ..recon_fit import Priors
from ..smpl import LandmarkRegressor, SMPL_Layer

_provider = None


def set_asset_provider(p):
    global _provider
    _provider = p


def get_asset_provider():
    global _provider
    if _provider is None:
        _provider = ReferenceAssets()
    return _provider


class AssetProvider:
    def smplh(self, gender: str, device) -> SMPL_Layer: raise NotImplementedError
    def body25(self, device) -> LandmarkRegressor: raise NotImplementedError
    def prior_arrays(self) -> Dict[str, np.ndarray]: raise NotImplementedError          # body / hand priors + part_labels (assets.npz keys)
    def sifnet(self, opt, device) -> CHORETriplaneVisibility: raise NotImplementedError
    def object_template(self, name: str) -> Tuple[np.ndarray, np.ndarray]: raise NotImplementedError   # centred (verts [V,3], faces [F,3])

    def priors(self, device) -> Priors:
        return Priors(self.prior_arrays(), device)


class SyntheticAssets(AssetProvider):
    """Seeded stand-ins: synthetic SMPL-H model (6890 vertices, 52 joints), random-init SIF-Net, a closed ellipsoid as the object template;
    the priors, regressors and part labels are the reference's own (exported to ``assets_npz`` by tests/golden/make_golden.py)."""

    def __init__(self, assets_npz: str, seed: int = 0):
        self.arrays = dict(np.load(assets_npz))
        self.seed = seed
        self._cache = {}

    def smplh(self, gender, device):
        from ..synth_smpl import synthetic_body_mesh, synthetic_smplh
        key = ("smplh", str(device))
        if key not in self._cache:
            model = synthetic_smplh(seed=3)
            layer = SMPL_Layer.from_buffers(model, model["parents"], device, gender=gender)
            _, faces = synthetic_body_mesh()             # connectivity with SMPL's vertex count (the synthetic model's own faces are random)
            layer.faces = faces.astype(np.int32)
            self._cache[key] = layer
        return self._cache[key]

    def body25(self, device):
        a = self.arrays
        return LandmarkRegressor(np.stack([a["body25_row"], a["body25_col"]]), a["body25_val"], a["body25_shape"], device)

    def prior_arrays(self):
        return self.arrays

    def sifnet(self, opt, device):
        from ..config import resolve_dims
        from ..synth import synthetic_state_dict
        key = ("sifnet", str(device))
        if key not in self._cache:
            net = CHORETriplaneVisibility(opt, device=device).eval()
            net.load_state_dict(synthetic_state_dict(resolve_dims(opt), seed=self.seed))
            self._cache[key] = net
        return self._cache[key]

    def object_template(self, name):
        from ..synth_smpl import synthetic_body_mesh
        v, f = synthetic_body_mesh(rings=20, segments=20, radii=(0.3, 0.25, 0.2))
        return (v - v.mean(0)).astype(np.float32), f.astype(np.int32)


class ReferenceAssets(AssetProvider):
    """The reference's files (paths from PATHS.yml / VT_* variables, see ``shims.paths``):
    SMPL_MODEL_ROOT/SMPLH_<gender>.pkl (lib_smpl/smplpytorch/smplpytorch/pytorch/smpl_layer.py:30-71), SMPL_ASSETS_ROOT/{body25_regressor,
    priors, smpl_parts_dense}.pkl, CODE/experiments/<exp_name>/checkpoints/*.tar (recon/gen/generator.py:36-50) and
    BEHAVE_ROOT/objects/<name>/<name>*.ply.  The model pickles hold chumpy arrays: they need chumpy importable, as in the reference."""

    def __init__(self):
        from . import paths
        self.p = paths()

    def _need(self, key):
        if key not in self.p:
            raise RuntimeError(f"{key} is not set: put it in PATHS.yml (as the reference does) or export VT_{key}")
        return self.p[key]

    def smplh(self, gender, device):
        f = os.path.join(self._need("SMPL_MODEL_ROOT"), f"SMPLH_{gender}.pkl")
        with open(f, "rb") as fh:
            d = pkl.load(fh, encoding="latin1")
        arr = lambda k: np.asarray(d[k].r if hasattr(d[k], "r") else d[k])
        J_reg = d["J_regressor"]
        J_reg = np.asarray(J_reg.todense()) if hasattr(J_reg, "todense") else np.asarray(J_reg)
        buffers = {"th_v_template": torch.from_numpy(arr("v_template")[None]), "th_shapedirs": torch.from_numpy(arr("shapedirs")[:, :, :10].copy()),
                   "th_posedirs": torch.from_numpy(arr("posedirs")), "th_J_regressor": torch.from_numpy(J_reg),
                   "th_weights": torch.from_numpy(arr("weights")), "th_faces": torch.from_numpy(arr("f").astype(np.int64))}
        parents = [int(x) for x in np.asarray(d["kintree_table"])[0].astype(np.int64)]
        parents[0] = -1
        return SMPL_Layer.from_buffers(buffers, parents, device, gender=gender)

    def prior_arrays(self):
        root = self._need("SMPL_ASSETS_ROOT")
        npz = os.path.join(root, "vistracker_b200_assets.npz")
        if os.path.isfile(npz):
            return dict(np.load(npz))
        raise RuntimeError(f"export the reference's assets once with tests/golden/make_golden.py --only assets and copy assets.npz to {npz}")

    def body25(self, device):
        a = self.prior_arrays()
        return LandmarkRegressor(np.stack([a["body25_row"], a["body25_col"]]), a["body25_val"], a["body25_shape"], device)

    def sifnet(self, opt, device):
        ck = getattr(opt, "checkpoint", None)
        folder = os.path.join(self.p.get("CODE", "."), "experiments", opt.exp_name, "checkpoints")
        files = sorted(glob.glob(os.path.join(folder, "*.tar")))
        if ck is not None:
            files = [f for f in files if ck in os.path.basename(f)] or [ck]
        if not files:
            raise RuntimeError(f"no checkpoint under {folder}")
        net = CHORETriplaneVisibility(opt, device=device).eval()
        ckpt = torch.load(files[-1], map_location="cpu")                       # recon/gen/generator.py:296-308
        net.load_state_dict(ckpt["model_state_dict"] if "model_state_dict" in ckpt else ckpt)
        return net

    def object_template(self, name):
        from ..io import load_ply
        files = sorted(glob.glob(os.path.join(self._need("BEHAVE_ROOT"), "objects", name, f"{name}*.ply")))
        if not files:
            raise RuntimeError(f"no template for object '{name}' under BEHAVE_ROOT/objects")
        v, f = load_ply(files[0])
        return (v - v.mean(0)).astype(np.float32), f.astype(np.int32)
