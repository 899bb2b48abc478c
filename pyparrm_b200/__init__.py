"""pyparrm_b200: PARRM's period search and comb filter on NVIDIA B200 (sm_100a).

Drop-in for the public surface of ``pyparrm`` (reference ``src/pyparrm/__init__.py:3-6``):
``PARRM``, ``get_example_data_paths`` and ``__version__``.  ``import pyparrm_b200 as pyparrm``
or :func:`install_as_pyparrm` lets existing callers run unchanged.
"""

__version__ = "1.2.0dev+b200.r1"

from .data import get_example_data_paths
from .parrm import PARRM
from ._engine import pinned_empty
from ._sharding import disable as disable_sharding
from ._sharding import enable as enable_sharding


def install_as_pyparrm() -> None:
    """Register this package under the name ``pyparrm`` so ``from pyparrm import PARRM`` works."""
    import sys

    sys.modules.setdefault("pyparrm", sys.modules[__name__])
    sys.modules.setdefault("pyparrm.data", sys.modules[__name__ + ".data"])
    sys.modules.setdefault("pyparrm.parrm", sys.modules[__name__ + ".parrm"])


__all__ = ["PARRM", "get_example_data_paths", "pinned_empty", "install_as_pyparrm",
           "enable_sharding", "disable_sharding", "__version__"]
