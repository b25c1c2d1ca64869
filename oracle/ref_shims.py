"""ORACLE SUPPORT — test infrastructure only.

Imports the *unmodified* reference modules from /root/reference on a CUDA-less host, with the
stubs listed in SURVEY.md §8c.  Only used in the build container (the reference tree does not
exist on the GPU box) by oracle/make_golden.py (golden vectors checked in tests/test_oracle_golden.py) and tests/test_reference_glue_cpu.py.
Nothing from the reference is copied; the stubs only satisfy imports the hot path never calls.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("VERSBAND_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ldm"))


def _stub(name: str, **attrs):
    if name in sys.modules:
        mod = sys.modules[name]
    else:
        mod = types.ModuleType(name)
        sys.modules[name] = mod
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


def install() -> None:
    """Make `import ldm...` resolve to the reference and stub the absent third-party imports."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
    try:
        import pytorch_lightning  # noqa: F401
    except Exception:
        pl = _stub("pytorch_lightning", LightningModule=torch.nn.Module)
        _stub("pytorch_lightning.utilities")
        _stub("pytorch_lightning.utilities.rank_zero", rank_zero_only=lambda f: f)
        _stub("pytorch_lightning.utilities.distributed", rank_zero_only=lambda f: f)
        pl.utilities = sys.modules["pytorch_lightning.utilities"]
    if not torch.cuda.is_available():
        # vocal2music_moe.py:469 calls .cuda() while building the RoPE table
        torch.Tensor.cuda = lambda self, *a, **k: self  # type: ignore[assignment]


def load_dit_module():
    install()
    return importlib.import_module("ldm.modules.diffusionmodules.vocal2music_moe")


def load_vae_module():
    install()
    return importlib.import_module("ldm.models.autoencoder1d")


def load_hifigan_module():
    """vocoder/hifigan/__init__.py pulls librosa; load the generator file directly."""
    install()
    path = os.path.join(REFERENCE_ROOT, "vocoder/hifigan/modules/hifigan.py")
    spec = importlib.util.spec_from_file_location("_ref_hifigan_modules", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
