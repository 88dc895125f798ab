#!/usr/bin/env python
"""Per-kernel histogram of the Blackwell-specific SASS opcodes in libmadeleine_b200.so (tcgen05 MMA = UTCHMMA, TMA loads =
UTMALDG, TMEM loads = LDTM, tcgen05.commit / mbarrier arrive = UTCBAR, bulk async copies = UBLKCP, mbarrier waits = SYNCS):
the evidence that the GEMMs run on the 5th-generation tensor cores through TMA and TMEM, checkable without rebuilding.

    python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "madeleine_b200", "libmadeleine_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMALDG.2CTA", "LDTM", "UTCBAR", "UTCBAR.2CTA", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "FFMA", "MUFU", "RED", "ATOMG"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.splitlines()
    names = iter(demangle)
    table, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next(names, m.group(1))
            cur = re.sub(r"\((int|bool|unsigned int)\)", "", cur)
            cur = re.sub(r"\([^()]*(\([^()]*\)[^()]*)*\)\s*$", "", cur)          # drop the parameter list, keep template arguments
            while cur in table:
                cur += "'"
            table[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        table[cur]["_total"] += 1
        base = op.split(".")[0]
        if base in ("UTCHMMA", "UTMALDG", "UTCBAR"):
            table[cur][base + (".2CTA" if ".2CTA" in op else "")] += 1
        elif base in OPS:
            table[cur][base] += 1
    print(f"# {os.path.relpath(LIB, REPO)}: SASS opcode counts per kernel (cuobjdump -sass, sm_100a)")
    print(f"# {'kernel':86s} " + " ".join(f"{o:>12s}" for o in OPS) + f" {'instructions':>12s}")
    tot = collections.Counter()
    for k, c in table.items():
        print(f"{k[:88]:88s} " + " ".join(f"{c.get(o, 0):12d}" for o in OPS) + f" {c['_total']:12d}")
        tot.update(c)
    print(f"{'TOTAL':88s} " + " ".join(f"{tot.get(o, 0):12d}" for o in OPS) + f" {tot['_total']:12d}")


if __name__ == "__main__":
    sys.exit(main())
