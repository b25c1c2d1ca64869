# namespace bridge: keep the reference's sub-modules importable next to the B200 overrides, and re-export what
# the reference package's own __init__ exports (vocoder/hifigan/__init__.py:1-2) so that the reference's call
# site `from vocoder.hifigan import HifiGAN, CodeUpsampleHifiGan` (scripts/test_final.py:21) keeps working.
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)

from versband_b200.modules import HifiGAN  # noqa: E402,F401  (B200 implementation)


def __getattr__(name):
    # classes outside the hot path stay the reference's own; resolved lazily (they pull librosa etc.)
    if name == "CodeUpsampleHifiGan":
        from .hifigan import CodeUpsampleHifiGan
        return CodeUpsampleHifiGan
    if name == "HifiGAN_NSF":
        import importlib
        return importlib.import_module(__name__ + ".hifigan_nsf").HifiGAN_NSF
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
