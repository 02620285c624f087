#!/usr/bin/env python
"""SASS opcode histogram of libpnr_sm100.so -> profiles/r2_sass_histogram.txt (the evidence that the hot kernels are
Blackwell-native: UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pixel-nerf_b200", "lib", "libpnr_sm100.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "UCGABAR", "HMMA", "FFMA",
       "F2FP", "MUFU", "RED", "ATOM", "LDG", "STG", "LDS", "STS")


def main():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    out = ["# SASS opcode histogram of pixel-nerf_b200/lib/libpnr_sm100.so (cuobjdump -sass, sm_100a), per kernel: the opcodes",
           "# that prove the Blackwell-native path (tcgen05 = UTCHMMA / UTCBAR, TMEM = LDTM / STTM, TMA bulk engine = UBLKCP,",
           "# mbarrier = SYNCS, cluster barrier = UCGABAR) and the 12 most frequent opcode families.",
           "# Regenerate: python scripts/sass_histogram.py", ""]
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        ops = collections.Counter()
        for line in f.split("\n"):
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
            if m:
                ops[m.group(1)] += 1
        tot = sum(ops.values())
        if tot < 50:
            continue
        fam = collections.Counter()
        for o, c in ops.items():
            fam[o.split(".")[0]] += c
        try:
            name = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0] or name
        except OSError:
            pass
        out.append(f"== {name}  ({tot} instructions)")
        out.append("   key: " + " ".join(f"{k}={fam[k]}" for k in KEY if fam.get(k)))
        detail = [f"{o}={c}" for o, c in sorted(ops.items()) if o.split(".")[0] in ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP")]
        if detail:
            out.append("   tcgen05 / TMA detail: " + " ".join(detail))
        out.append("   top: " + " ".join(f"{o}={c}" for o, c in fam.most_common(12)))
    dst = os.path.join(ROOT, "profiles", "r2_sass_histogram.txt")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(l for l in out if "tcgen05" in l or l.startswith("==")))


if __name__ == "__main__":
    sys.exit(main())
