"""Entry-point shims: the reference's command-line programs of the per-frame hot path under their own module names and argparse flags, running on
the B200 operators of this package, so that ``scripts/demo.sh`` steps 1, 2b, 3, 4 and 6 keep their command lines:

    python preprocess/fit_SMPLH_30fps.py -s SEQ -bs 512                                  -> shims.fit_SMPLH_30fps.main
    python preprocess/fit_SMPLH_smoothed.py -sn smplt-smoothed -s SEQ                    -> shims.fit_SMPLH_smoothed.main
    python render/render_triplane_nr.py -s SEQ                                            -> shims.render_triplane_nr.main
    python recon/recon_fit_trivis_full.py tri-vis-l2 -sn ... -or ... -sr ... -s SEQ       -> shims.recon_fit_trivis_full.recon_fit

``install()`` registers the shim modules in ``sys.modules`` as ``preprocess.fit_SMPLH_30fps``, ``preprocess.fit_SMPLH_smoothed``,
``render.render_triplane_nr`` and ``recon.recon_fit_trivis_full`` (the reference's import paths); each module is also runnable with
``python -m vistracker_b200.shims.<name> <the reference's arguments>``.

What the reference reads from disk around these programs -- the BEHAVE sequence layout through its vendored ``behave`` toolkit, the model
pickles, checkpoints and object templates through PATHS.yml -- comes from two small layers: ``seqio`` (file names, cited line by line) and
``assets`` (an AssetProvider: the reference's folders by default, a seeded synthetic set for the tests).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_NAMES = {"preprocess.fit_SMPLH_30fps": "fit_SMPLH_30fps", "preprocess.fit_SMPLH_smoothed": "fit_SMPLH_smoothed",
          "render.render_triplane_nr": "render_triplane_nr", "recon.recon_fit_trivis_full": "recon_fit_trivis_full"}


def install(force: bool = False):
    """Make ``import recon.recon_fit_trivis_full`` (etc.) resolve to the shims.  Existing modules of those names (a checkout of the reference on
    sys.path that was already imported) are left alone unless ``force``."""
    done = []
    for ref_name, shim in _NAMES.items():
        pkg = ref_name.split(".")[0]
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
        if ref_name in sys.modules and not force:
            continue
        mod = importlib.import_module(f"{__name__}.{shim}")
        sys.modules[ref_name] = mod
        setattr(sys.modules[pkg], ref_name.split(".")[1], mod)
        done.append(ref_name)
    return done


def paths() -> dict:
    """RECON_PATH / GT_PACKED / ... as the reference resolves them: ``PATHS.yml`` in the working directory (lib_smpl/const.py, recon/*: read
    relative to cwd), overridable by environment variables VT_<KEY>; RECON_PATH falls back to ./recon_out."""
    d = {}
    if os.path.isfile("PATHS.yml"):
        import yaml
        with open("PATHS.yml") as f:
            d = dict(yaml.safe_load(f) or {})
    for k in ("RECON_PATH", "GT_PACKED", "SMPL_MODEL_ROOT", "SMPL_ASSETS_ROOT", "BEHAVE_ROOT", "CODE"):
        if os.environ.get("VT_" + k):
            d[k] = os.environ["VT_" + k]
    d.setdefault("RECON_PATH", os.path.abspath("recon_out"))
    return d
