"""Import the UNMODIFIED reference: from /root/reference in the build container, else from
the verbatim copy ``oracle/_ref/pyparrm`` made by ``oracle/vendor_ref.py`` (git-ignored; it
travels to the GPU box, where /root/reference does not exist).

TEST / BENCH INFRASTRUCTURE.  Used by ``oracle/make_golden.py`` (records reference outputs
into ``tests/golden/``), by CPU tests that skip when neither location exists, and by the CPU
legs of ``bench.py`` (``--impl reference`` and ``cpu_baseline``), which time the reference's
own methods.  Nothing in ``pyparrm_b200/`` imports this.

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

_MOUNTED = "/root/reference/src"
_VENDORED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _source_dir() -> str | None:
    for base in (_MOUNTED, _VENDORED):
        if os.path.isfile(os.path.join(base, "pyparrm", "parrm.py")):
            return base
    return None


REFERENCE_SRC = _source_dir() or _MOUNTED


def reference_available() -> bool:
    return _source_dir() is not None


def reference_location() -> str:
    """Where the reference would be imported from ("" when it is nowhere)."""
    return _source_dir() or ""


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
    src = _source_dir()
    if src is None:
        raise ImportError(f"reference neither mounted at {_MOUNTED} nor vendored at {_VENDORED} "
                          "(python oracle/vendor_ref.py)")
    _install_stand_ins()
    if src not in sys.path:
        sys.path.insert(0, src)
    import pyparrm

    if not os.path.abspath(pyparrm.__file__).startswith(src):
        raise ImportError(f"'pyparrm' resolved to {pyparrm.__file__}, not the reference")
    return pyparrm
