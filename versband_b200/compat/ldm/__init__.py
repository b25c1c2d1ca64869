# namespace bridge: keep the reference's sub-modules importable next to the B200 overrides
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
