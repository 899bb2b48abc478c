"""Development aid: compile filter_comb_e.cuh offline (nvcc, no GPU) for a named tap set with
the shape the library would pick, print registers / spills and the SASS opcode histogram.
python scripts/comb_e_sass.py cfg3 [steps_per_chunk prefetch ctas] [--sass out.sass]"""
import ctypes
import os
import subprocess
import sys
from collections import Counter

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import parrm_oracle as oracle  # noqa: E402
from pyparrm_b200 import _native as K  # noqa: E402

CASES = {
    "cfg2": (2000 / 130 * (1 + 3e-6), None, 2000, 0, "both"),
    "cfg3": (1000 / 145 * (1 + 3e-6), None, 2469, 0, "both"),
    "cfg4": (30000 / 130 * (1 + 3e-6), None, 2311, 0, "past"),
    "cfg4f": (30000 / 130 * (1 + 3e-6), None, 2311, 0, "future"),
    "cfg1": (1.3311148014466094, 0.01, 2000, 20, "both"),
}


def flags_for(name, dtype=K.F64, tuning=()):
    period, phw, hw, omit, direction = CASES[name]
    taps = oracle.tap_offsets(period, period / 50 if phw is None else phw, hw, omit, direction)
    plan, desc = K.plan_filter(taps, dtype)
    opts = K.FilterOptions()
    for key, val in zip(("steps_per_chunk", "prefetch_chunks", "ctas_per_sm"), tuning):
        setattr(opts, key, int(val))
    shape = np.zeros(12, dtype=np.int32)
    nbytes = ctypes.c_size_t(0)
    K.check(K.lib.parrm_filter_specialise_check(plan.ctypes.data, dtype, ctypes.byref(opts),
                                                shape.ctypes.data, ctypes.byref(nbytes)),
            "specialise_check")
    d, nk, m0, m1, nb0, nb1, _, u, pf, ctas, smem, threads = (int(v) for v in shape)
    boxes = desc["boxes"]
    wins = desc["windows"]
    order = [0, 1][:nk]
    if nk == 2 and wins[1] > wins[0]:
        order = [1, 0]
    off = [list(map(int, boxes[i])) for i in order] + [[0]]
    terms = [int(v) for b in boxes for v in b] + list(map(int, desc["plus"])) + list(map(int, desc["minus"]))
    back, fwd = max(0, max(terms)), max(0, -min(terms))
    j = lambda v: ",".join(str(int(x)) for x in v) if len(v) else "0"  # noqa: E731
    es_t = "double" if dtype == K.F64 else "float"
    D = dict(PE_T=es_t, PE_D=d, PE_NK=nk, PE_M0=m0, PE_M1=m1, PE_NB0=nb0, PE_NB1=nb1,
             PE_OFF0=j(off[0]), PE_OFF1=j(off[1] if nk == 2 else []), PE_NPLUS=len(desc["plus"]),
             PE_PLUS=j(desc["plus"]), PE_NMINUS=len(desc["minus"]), PE_MINUS=j(desc["minus"]),
             PE_CENTRE=desc["centre"], PE_U=u, PE_PF=pf, PE_NTAPS=desc["n_taps"],
             PE_WLO=min(int(taps[0]), 0), PE_WHI=max(int(taps[-1]), 0), PE_BACK=back, PE_FWD=fwd,
             PE_CTAS=ctas)
    if dtype == K.F32:
        D["PE_TMAX"] = "3.402823466e38f"
    print(name, "shape:", dict(d=d, nk=nk, m=(m0, m1), nb=(nb0, nb1), singles=len(desc["plus"]) + len(desc["minus"]),
                               u=u, pf=pf, ctas=ctas, smem=smem, threads=threads), "cubin", nbytes.value)
    return "".join(f"#define {k} {v}\n" for k, v in D.items())


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--") and not a.endswith(".sass")]
    name = args[0] if args else "cfg2"
    dtype = K.F32 if "--f32" in sys.argv else K.F64
    flags = flags_for(name, dtype, [a for a in args[1:4] if a.isdigit()])
    src = f"/tmp/comb_e_{name}.cu"
    with open(src, "w") as f:
        f.write(flags + '#include "filter_comb_e.cuh"\n')
    obj = f"/tmp/comb_e_{name}.o"
    cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
           "-lineinfo", "-Xptxas", "-v", "-I", os.path.join(ROOT, "pyparrm_b200/csrc"), "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    print("\n".join(l for l in r.stderr.splitlines() if "spill" in l or "Used" in l or "error" in l))
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    ops = Counter()
    for line in sass.splitlines():
        parts = line.split()
        if len(parts) > 1 and parts[0].startswith("/*") and parts[0].endswith("*/") and len(parts[0]) == 8:
            op = parts[2] if parts[1].startswith("@") else parts[1]
            ops[op.split(".")[0].rstrip(";")] += 1
    print("static SASS:", sum(ops.values()), dict(ops.most_common(14)))
    if "--sass" in sys.argv:
        open(sys.argv[sys.argv.index("--sass") + 1], "w").write(sass)
