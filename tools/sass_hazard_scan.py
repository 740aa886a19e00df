#!/usr/bin/env python3
"""Scan the SASS of a built library for one scheduling pattern that bit us on sm_100a (ptxas 12.9.86):

    CS2R Rd, SRZ ;            // zero the pair Rd:Rd+1
    @!P  IMAD.WIDE.U32 Rd, .. // predicated off at run time
    LEA  .., Rd, ..           // 5 issue cycles after the CS2R: observed to read the OLD content of Rd

The consumer 9 cycles away (Rd+1) saw the zero.  The pattern came from a 64-bit accumulator fed by predicated
byte loads (load_be64 in csrc/cdf_kernels.cu, since rewritten); it corrupted the decoder's bit window for
streams shorter than 22 bytes.  This script flags every reader of a CS2R destination closer than `--min`
issue cycles (default 7) inside straight-line code, so a rebuild that reintroduces the pattern fails CI.

usage: sass_hazard_scan.py lib.so [--min 7]   (exit status 1 when something is flagged)
"""
import re
import subprocess
import sys


def parse(so):
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout.split("\n")
    funcs, func, i = {}, None, 0
    while i < len(txt):
        m = re.search(r"Function : (\S+)", txt[i])
        if m:
            func = m.group(1)
            funcs[func] = []
        else:
            m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/", txt[i])
            m2 = re.match(r"\s*/\* 0x([0-9a-f]{16}) \*/", txt[i + 1]) if m and i + 1 < len(txt) else None
            if m2:
                stall = (int(m2.group(1), 16) >> 41) & 0xF  # issue-stall field of the control word
                funcs[func].append((int(m.group(1), 16), m.group(2).strip(), stall))
                i += 1
        i += 1
    return funcs


def regs(s):
    return [int(x) for x in re.findall(r"\bR(\d+)\b", s)]


def scan(so, min_cycles=7):
    hits = []
    for f, ins in parse(so).items():
        for k, (addr, text, stall) in enumerate(ins):
            m = re.match(r"(@!?U?P\d+\s+)?CS2R(\.32)?\s+R(\d+),", text)
            if not m:
                continue
            rd = int(m.group(3))
            live = {rd} if m.group(2) else {rd, rd + 1}
            cyc, j = stall, k + 1
            while j < len(ins) and cyc < min_cycles and live:
                a2, t2, st2 = ins[j]
                body = re.sub(r"^@!?U?P\d+\s+", "", t2)
                op = body.split()[0]
                if op.startswith(("BRA", "BSYNC", "BSSY", "EXIT", "RET", "CALL", "JMP")):
                    break
                ops = body.split(None, 1)[1] if " " in body else ""
                toks = [x.strip() for x in ops.split(",")]
                has_dst = not op.startswith(("ST", "RED", "ATOM", "BAR", "MEMBAR", "FENCE", "SYNCS", "UTMA"))
                dst = regs(toks[0]) if toks and has_dst else []
                srcs = set()
                for tk in (toks[1:] if dst else toks):
                    srcs |= set(regs(tk))
                if srcs & live:
                    hits.append((f, addr, text, a2, t2, cyc))
                    break
                if dst and not t2.startswith("@"):
                    live -= set(dst)
                    if "WIDE" in op or ".64" in op or op.startswith("CS2R"):
                        live -= {dst[0] + 1}
                cyc += st2
                j += 1
    return hits


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    mc = int(sys.argv[sys.argv.index("--min") + 1]) if "--min" in sys.argv else 7
    if "--min" in sys.argv:
        args = [a for a in args if a != str(mc)]
    h = scan(args[0], mc)
    for f, a, t, a2, t2, c in h:
        print(f"{f[:70]}  {a:05x} {t}  ->  {a2:05x} [{c} cycles] {t2}")
    print("flagged:", len(h))
    sys.exit(1 if h else 0)
