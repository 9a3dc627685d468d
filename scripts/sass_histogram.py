#!/usr/bin/env python
"""Opcode histogram per kernel of the shipped library (cuobjdump -sass), written to
profiles/sass_libsvihmm.txt: the evidence of which hardware paths each kernel uses
(UTC*MMA / LDTM = tcgen05 + TMEM, UTMALDG / UBLKCP = TMA, HMMA = legacy mma.sync, DFMA = FP64 pipe,
LDGSTS = cp.async, ...)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pysvihmm_b200", "lib", "libsvihmm.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF",
        "HMMA", "DFMA", "DADD", "DMUL", "FFMA", "FFMA2", "MUFU", "LDGSTS", "LDG", "STG", "LDS", "STS", "SHFL",
        "REDUX", "SYNCS", "ATOMG", "RED", "ATOMS", "BAR", "NANOSLEEP", "F2F", "I2F"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kern, hist, total = None, collections.OrderedDict(), {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(.*$", "", kern)
            hist[kern] = collections.Counter(); total[kern] = 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and kern:
            op, mods = m.group(1), m.group(2)
            total[kern] += 1
            key = op
            if op == "FFMA2" or (op == "FFMA" and False):
                key = "FFMA2"
            if op in ("HMMA", "LDTM", "UTCHMMA"):
                key = op
            hist[kern][key] += 1
    lines = ["# cuobjdump -sass %s : static instruction counts per kernel (selected opcodes)" % os.path.relpath(LIB, ROOT),
             "# %d kernels" % len(hist), ""]
    allc = collections.Counter()
    for k, h in hist.items():
        allc.update(h)
        sel = ["%s=%d" % (key, h[key]) for key in KEYS if h.get(key)]
        lines.append("%s\n    total=%d  %s" % (k, total[k], "  ".join(sel)))
    lines += ["", "# whole library: " + "  ".join("%s=%d" % (key, allc[key]) for key in KEYS if allc.get(key))]
    dst = os.path.join(ROOT, "profiles", "sass_libsvihmm.txt")
    open(dst, "w").write("\n".join(lines) + "\n")
    print(lines[-1])


if __name__ == "__main__":
    sys.exit(main())
