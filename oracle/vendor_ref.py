"""Recipe: place an UNMODIFIED copy of the reference package under ``oracle/_ref/``.

TEST / BENCH INFRASTRUCTURE.  The reference (neuromodulation/PyPARRM) is pure Python; there
is nothing to compile.  ``pip install`` of it fails in this image (its build backend
``hatchling`` is absent and there is no index), so "installing" it means copying the package
directory ``/root/reference/src/pyparrm`` verbatim:

    python oracle/vendor_ref.py        # -> oracle/_ref/pyparrm/   (+ MANIFEST.sha256)

``oracle/_ref/`` is git-ignored (reference sources never enter this repository's history)
but is NOT gpurun-ignored, so the copy travels to the GPU box, where ``/root/reference`` does
not exist.  ``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` legs import it
through ``oracle/ref_shim.py`` (``pqdm`` / ``matplotlib`` stand-ins, dispatch only) and time
the reference's own ``PARRM`` methods.  ``__graft_entry__.build()`` runs this recipe whenever
``/root/reference`` is present.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/pyparrm"
DST = os.path.join(HERE, "_ref", "pyparrm")


def vendor(verbose: bool = True) -> bool:
    """Copy the package; returns False when the reference is not mounted."""
    if not os.path.isdir(SRC):
        if verbose:
            print(f"vendor_ref: {SRC} not present; keeping {DST} as it is")
        return False
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    lines = []
    for base, _, files in sorted(os.walk(DST)):
        for name in sorted(files):
            path = os.path.join(base, name)
            with open(path, "rb") as fh:
                digest = hashlib.sha256(fh.read()).hexdigest()
            lines.append(f"{digest}  {os.path.relpath(path, DST)}")
    with open(os.path.join(HERE, "_ref", "MANIFEST.sha256"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    if verbose:
        print(f"vendor_ref: copied {len(lines)} files to {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if vendor() else 1)
