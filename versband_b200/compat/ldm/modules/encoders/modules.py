"""B200 override of the YAML target `ldm.modules.encoders.modules.FrozenTextVocalEmbedder`
(configs/vocal2music.yaml:70-74). Every other name of the reference file (CLAP / CLIP / BERT embedders: not
on the path) is forwarded, lazily, to the reference's own module further down the extended package path."""
from versband_b200.modules import FrozenTextVocalEmbedder  # noqa: F401


def _reference_module():
    import importlib.util
    import os
    import sys

    name = __name__.rsplit(".", 1)[0] + "._reference_modules"
    if name in sys.modules:
        return sys.modules[name]
    pkg = sys.modules[__name__.rsplit(".", 1)[0]]
    here = os.path.dirname(os.path.abspath(__file__))
    for d in pkg.__path__:
        f = os.path.join(d, "modules.py")
        if os.path.abspath(d) != here and os.path.exists(f):
            spec = importlib.util.spec_from_file_location(name, f)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
            return mod
    raise ImportError("the reference's ldm/modules/encoders/modules.py is not on sys.path behind versband_b200/compat")


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return getattr(_reference_module(), name)
