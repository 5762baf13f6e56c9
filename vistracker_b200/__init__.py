"""vistracker_b200 -- B200-native (sm_100a) implementation of the VisTracker per-frame hot path.

Public surface mirrors the reference objects for this path:
  * ``CHORETriplaneVisibility`` -- SIF-Net ``filter`` / ``query`` / ``get_preds`` (model/chore_tri_vis.py)
  * ``load_configs`` / ``default_options`` -- experiment options (config/config_loader.py)
The arithmetic lives in ``libvistracker_sm100a.so`` (C ABI: include/vistracker_b200.h).
"""
from .config import default_options, load_configs, resolve_dims  # noqa: F401
from .sifnet import CHORETriplaneVisibility  # noqa: F401

__all__ = ["CHORETriplaneVisibility", "default_options", "load_configs", "resolve_dims"]
