"""SASS mnemonic histogram per kernel of the in-tree libgode.so (cuobjdump -sass): the instructions that prove which unit a
kernel runs on — tcgen05 (UTC*MMA, LDTM / STTM = TMEM loads / stores), bulk copies (UBLKCP / UTMALDG), MUFU.TANH, packed
FP32 FMA (FFMA2), shared-memory traffic, shuffles, grid-dependency control.  Writes profiles/r2_sass_hist.txt.
    python scripts/sass_hist.py
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gan_ode_b200", "csrc", "libgode.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "MUFU.TANH", "MUFU.EX2",
         "MUFU.RCP", "MUFU.SIN", "MUFU.LG2", "FFMA2", "FFMA", "DFMA", "LDS", "STS", "SHFL", "LDG", "STG", "RED", "ATOM",
         "BAR.SYNC", "SYNCS", "ACQBULK", "PREEXIT", "ELECT"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(.*$", "", cur).replace("gode::", "")
            kernels[cur] = collections.Counter()
            continue
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    kernels[cur][w] += 1
    cols = [w for w in WATCH if any(c[w] for c in kernels.values())]
    lines = ["SASS mnemonic counts per kernel of gan_ode_b200/csrc/libgode.so (cuobjdump -sass, sm_100a); static counts, not executed counts.",
             "{:<78s} {:>7s} ".format("kernel", "total") + " ".join("{:>9s}".format(c) for c in cols)]
    for k, c in kernels.items():
        lines.append("{:<78s} {:>7d} ".format(k[:78], c["total"]) + " ".join("{:>9d}".format(c[w]) for w in cols))
    text = "\n".join(lines) + "\n"
    path = os.path.join(ROOT, "profiles", "r2_sass_hist.txt")
    open(path, "w").write(text)
    sys.stdout.write(text[:3000])


if __name__ == "__main__":
    main()
