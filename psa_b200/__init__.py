"""psa_b200: B200-native spectral-energy-density (SED) hot path behind PSA's SEDCalculator API."""
from .trajectory import Trajectory
from .sed import SED
from .directions import parse_direction

__all__ = ["Trajectory", "SED", "parse_direction", "SEDCalculator", "iSEDReconstructor"]
__version__ = "0.1.0"


def __getattr__(name):
    # the calculator pulls in torch + the CUDA library; keep `import psa_b200` light
    if name in ("SEDCalculator", "iSEDReconstructor"):
        from . import calculator
        return getattr(calculator, name)
    raise AttributeError(name)
