"""SASS opcode histogram per kernel of the built library (cuobjdump -sass): which Blackwell paths each kernel uses.
python tools/sass_histogram.py [lib.so] > profiles/r2/sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "lac_b200", "_lib", "liblac_b200.so")
text = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", text)))
kernels = collections.OrderedDict()
cur = None
for line in text.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = kernels.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur is not None:
        cur[m.group(1).split(".")[0] if not m.group(1).startswith(("UBLKCP", "SYNCS", "REDUX", "FFMA2", "FADD2", "STAS", "UTMA"))
            else m.group(1)] += 1
print(f"{os.path.basename(so)}: cubins for {', '.join(arch)}; {len(kernels)} kernels")
notable = ("UBLKCP", "UTMALDG", "SYNCS", "STAS", "FFMA2", "FADD2", "REDUX", "MUFU", "F2I", "I2F", "DFMA", "DMUL", "HMMA", "UTC")
for name, c in kernels.items():
    total = sum(c.values())
    marks = {k: v for k, v in c.items() if k.startswith(notable)}
    top = ", ".join(f"{k} {v}" for k, v in c.most_common(8))
    print(f"\n{name}\n  {total} instructions; top: {top}")
    if marks:
        print("  notable: " + ", ".join(f"{k} {v}" for k, v in sorted(marks.items())))
