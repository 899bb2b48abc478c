"""Import the UNMODIFIED reference from /root/reference -- build container only.

TEST INFRASTRUCTURE.  ``/root/reference`` does not exist on the GPU box, so
nothing under ``tests/ -m gpu``, ``smoke()`` or ``bench.py`` may call this; it
is used by ``oracle/make_golden.py`` (which records reference outputs into
``tests/golden/``) and by CPU tests that skip when the path is absent.

The reference's ``import pyparrm`` fails here because ``pqdm`` and
``matplotlib`` are not installed (``parrm.py:9``, ``_utils/_plotting.py:10-11``)
and cannot be (no network).  Two in-memory stand-ins make it importable without
touching its arithmetic:

* ``pqdm.threads.pqdm(array, function, n_jobs, argument_type="kwargs", ...)`` --
  an order-preserving thread map of ``function(**item)``; the two call sites
  (``parrm.py:445-454``, ``:510-517``) map pure functions, so results are
  unchanged.
* empty ``matplotlib`` / ``matplotlib.pyplot`` / ``matplotlib.widgets`` modules
  (only the GUI explorer uses them).
"""

from __future__ import annotations

import os
import sys
import types
from concurrent.futures import ThreadPoolExecutor

REFERENCE_SRC = "/root/reference/src"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "pyparrm"))


def _thread_map(array, function, n_jobs, argument_type=None, **_ignored):
    def call(item):
        if argument_type == "kwargs":
            return function(**item)
        if argument_type == "args":
            return function(*item)
        return function(item)

    items = list(array)
    if n_jobs <= 1:
        return [call(item) for item in items]
    with ThreadPoolExecutor(max_workers=n_jobs) as pool:
        return list(pool.map(call, items))


def _install_stand_ins() -> None:
    if "pqdm" not in sys.modules:
        pqdm_pkg = types.ModuleType("pqdm")
        pqdm_threads = types.ModuleType("pqdm.threads")
        pqdm_threads.pqdm = _thread_map
        pqdm_pkg.threads = pqdm_threads
        sys.modules["pqdm"] = pqdm_pkg
        sys.modules["pqdm.threads"] = pqdm_threads
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = types.ModuleType("matplotlib")
        pyplot = types.ModuleType("matplotlib.pyplot")
        widgets = types.ModuleType("matplotlib.widgets")
        widgets.RadioButtons = type("RadioButtons", (), {})
        widgets.TextBox = type("TextBox", (), {})
        mpl.pyplot = pyplot
        mpl.widgets = widgets
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = pyplot
        sys.modules["matplotlib.widgets"] = widgets


def import_reference():
    """Return the reference ``pyparrm`` package (raises if it is not mounted)."""
    if not reference_available():
        raise ImportError(f"reference not mounted at {REFERENCE_SRC}")
    _install_stand_ins()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import pyparrm

    if not os.path.abspath(pyparrm.__file__).startswith(REFERENCE_SRC):
        raise ImportError(f"'pyparrm' resolved to {pyparrm.__file__}, not the reference")
    return pyparrm
