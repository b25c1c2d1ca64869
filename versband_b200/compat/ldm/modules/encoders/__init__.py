# namespace bridge: the reference's other encoder sub-modules stay importable
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
