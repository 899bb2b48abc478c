"""Static SASS opcode histograms of the shipped library (cuobjdump -sass), written in the
format of profiles/r2/sass_opcodes.txt: python scripts/sass_opcodes.py [lib.so] > out.txt"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "pyparrm_b200", "libparrm_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)),
                       capture_output=True, text=True).stdout.splitlines()
kernels, cur = OrderedDict(), None
it = iter(names)
for line in sass.splitlines():
    if "Function :" in line:
        cur = kernels.setdefault(next(it), Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
KEY = ["UBLKCP", "SYNCS", "DMMA", "LDGSTS", "UTMALDG", "UTMASTG", "UTCHMMA", "UTCQMMA", "LDTM", "STTM",
       "DFMA", "DADD", "LDS", "STS", "LDG", "STG", "BAR", "SHFL"]
total = Counter()
for c in kernels.values():
    total.update(c)
print(f"== {os.path.basename(lib)} (pre-built kernels): {len(kernels)} kernels, "
      f"{sum(total.values())} instructions")
print("  key opcodes: " + ", ".join(f"{k}={total[k]}" for k in KEY))
for name, c in sorted(kernels.items(), key=lambda kv: -sum(kv[1].values())):
    print(f"  {sum(c.values()):6d}  {name[:110]}")
    print("          top: " + ", ".join(f"{k}:{v}" for k, v in c.most_common(8)))
    special = [f"{k}:{c[k]}" for k in ("UBLKCP", "SYNCS", "DMMA", "LDGSTS") if c[k]]
    if special:
        print("          async/tensor: " + ", ".join(special))
