"""versband_b200 — Blackwell-native (sm_100a) AccompBand inference hot path.

Hand-written CUDA kernels behind a C-ABI library (include/versband.h), driven by PyTorch
modules that mirror the reference's nn.Module surfaces. No fallback path exists.
"""
__version__ = "0.1.0"


def install_compat() -> str:
    """Put the shadow `ldm` / `vocoder` packages first on sys.path so that the reference's YAML
    `target:` strings instantiate the B200 modules. Returns the inserted path."""
    import os
    import sys

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "compat")
    if path in sys.path:
        sys.path.remove(path)
    sys.path.insert(0, path)
    # shadow packages imported BEFORE the reference checkout was put on sys.path extended their __path__ over
    # what existed then: refresh them so `ldm.models.diffusion...`, `vocoder.hifigan.modules...` keep resolving
    from pkgutil import extend_path
    for name in ("ldm", "ldm.models", "ldm.modules", "ldm.modules.diffusionmodules", "ldm.modules.encoders", "vocoder",
                 "vocoder.hifigan"):
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, "__path__"):
            mod.__path__ = extend_path(list(mod.__path__), name)
    return path


def instantiate_from_config(config: dict):
    """`{"target": "pkg.mod.Class", "params": {...}}` -> object (reference ldm/util.py:110-125)."""
    import importlib

    module, cls = config["target"].rsplit(".", 1)
    return getattr(importlib.import_module(module), cls)(**dict(config.get("params", {}) or {}))
