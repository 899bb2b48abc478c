"""pyparrm_b200: PARRM's period search and comb filter on NVIDIA B200 (sm_100a).

Drop-in for the public surface of ``pyparrm`` (reference ``src/pyparrm/__init__.py:3-6``):
``PARRM``, ``get_example_data_paths`` and ``__version__``.  ``import pyparrm_b200 as pyparrm``
or :func:`install_as_pyparrm` lets existing callers run unchanged.
"""

__version__ = "1.2.0dev+b200.r1"

from .data import get_example_data_paths

# Everything that needs libparrm_b200.so is resolved on first use (PEP 562), so helpers with no
# device work -- ``pyparrm_b200.synthetic``, ``pyparrm_b200.data`` -- can be imported without
# mapping the library (bench.py's reference arm relies on that).  Accessing ``PARRM`` on a
# machine where the library is not built raises ImportError: there is no CPU fallback.
_LAZY = {
    "PARRM": (".parrm", "PARRM"),
    "pinned_empty": ("._engine", "pinned_empty"),
    "pin_array": ("._engine", "pin_array"),
    "enable_sharding": ("._sharding", "enable"),
    "disable_sharding": ("._sharding", "disable"),
}


def __getattr__(name):
    if name in _LAZY:
        import importlib

        module, attr = _LAZY[name]
        value = getattr(importlib.import_module(module, __name__), attr)
        globals()[name] = value
        return value
    if name in ("parrm", "_engine", "_native", "_sharding", "_neldermead", "_utils"):
        import importlib

        return importlib.import_module("." + name, __name__)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def install_as_pyparrm() -> None:
    """Register this package under the name ``pyparrm`` so ``from pyparrm import PARRM`` works."""
    import sys

    import importlib

    sys.modules.setdefault("pyparrm", sys.modules[__name__])
    for sub in ("data", "parrm", "_utils", "_utils._power"):
        sys.modules.setdefault("pyparrm." + sub, importlib.import_module("." + sub, __name__))


__all__ = ["PARRM", "get_example_data_paths", "pinned_empty", "pin_array", "install_as_pyparrm",
           "enable_sharding", "disable_sharding", "__version__"]
