"""SASS evidence that the shipped library is Blackwell-native (profiles/rNN_sass_*.txt): per kernel, the count of
tcgen05 / TMEM / TMA mnemonics (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UBLKCP,
legacy mma.sync -> HMMA), plus the instruction lines themselves for the update GEMM.
    python tools/sass_evidence.py > profiles/r02_sass_k_gemm512.txt"""
import collections
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from buckgnn_b200 import build

MN = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "HMMA", "SYNCS", "USETMAXREG", "LDGSTS")
so = build.LIB_PATH
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
print(f"# cuobjdump -sass {os.path.relpath(so)}   ({os.path.getsize(so)} bytes, sm_100a)")
fn, per, lines = None, collections.OrderedDict(), {}
for ln in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        fn = m.group(1)
        per[fn] = collections.Counter()
        lines[fn] = []
        continue
    if fn is None:
        continue
    for k in MN:
        if re.search(r"\b" + k + r"\b|\b" + k + r"\.", ln):
            per[fn][k] += 1
            if k in ("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR"):
                mm = re.match(r"\s*/\*([0-9a-f]+)\*/\s*(.*?;)", ln)
                lines[fn].append(f"/*{mm.group(1)}*/ {mm.group(2)}" if mm else ln.strip())
demangle = subprocess.run(["c++filt"] + list(per), capture_output=True, text=True).stdout.splitlines()
names = dict(zip(per, demangle))
print("\n## mnemonic counts per kernel (kernels with none of them omitted)")
tot = collections.Counter()
for fn, c in per.items():
    if not c:
        continue
    tot.update(c)
    print(f"{names[fn][:110]:110s} " + " ".join(f"{k}={v}" for k, v in c.items()))
print("\n## totals: " + " ".join(f"{k}={v}" for k, v in tot.items()))
target = [f for f in per if "k_gemm512ILi2E6__halfLi1ELb0" in f]
if target:
    f = target[0]
    print(f"\n## {names[f]}: every tcgen05.mma / tcgen05.ld / TMA instruction (address, SASS)")
    for ln in lines[f]:
        print("   ", ln)
