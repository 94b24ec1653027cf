#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific SASS mnemonics in the shipped library (cuobjdump -sass): tcgen05 MMA
(UTCHMMA / UTCQMMA ...), TMEM loads (LDTM), TMEM alloc (UTCATOMSWS / UTCALLOC...), TMA loads/stores (UTMALDG /
UTMASTG), mbarrier (SYNCS), cluster barriers (UTCBAR).  Evidence that the kernels are tcgen05/TMEM/TMA code.

    python scripts/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "deqsci_b200", "libdeqsci.so")
PAT = re.compile(r"\b(UTC[A-Z0-9]*MMA[.A-Z0-9_]*|LDTM[.A-Z0-9_]*|STTM[.A-Z0-9_]*|UTMALDG[.A-Z0-9_]*|UTMASTG[.A-Z0-9_]*|UTMAPF[.A-Z0-9_]*|"
                 r"UTCBAR[.A-Z0-9_]*|UTCATOMSWS[.A-Z0-9_]*|UTCCP[.A-Z0-9_]*|SYNCS[.A-Z0-9_]*|UBLKCP[.A-Z0-9_]*|ACQBULK|"
                 r"HMMA[.A-Z0-9_]*|ELECT|UCGABAR_[A-Z]+|FFMA|LDG[.A-Z0-9_]*|STG[.A-Z0-9_]*)\b")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None or "/*" not in line:
            continue
        m = PAT.search(line)
        if m:
            op = m.group(1).rstrip(".")
            base = op.split(".")[0]
            cur[op if base.startswith(("UTC", "UTMA", "LDTM", "STTM")) else base] += 1
        cur["_instructions"] += 1 if re.search(r"/\*[0-9a-f]{4}\*/", line) else 0
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels) + "\n", capture_output=True, text=True).stdout.splitlines()
    print("# cuobjdump -sass deqsci_b200/libdeqsci.so (sm_100a): Blackwell mnemonics per kernel")
    print("# UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG = TMA tensor load/store,")
    print("# UTCBAR = tcgen05.commit (mbarrier arrive, .2CTA.MULTICAST for CTA pairs), SYNCS = mbarrier ops, UBLKCP = cp.async.bulk\n")
    tot = collections.Counter()
    for (name, c), dn in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dn)
        keys = [k for k in sorted(c) if k != "_instructions" and k not in ("FFMA", "LDG", "STG")]
        print("%s  [%d instructions]" % (short, c["_instructions"]))
        print("    " + (", ".join("%s x%d" % (k, c[k]) for k in keys) if keys else "(no tensor / TMA instructions)")
              + "   | FFMA x%d LDG x%d STG x%d" % (c["FFMA"], c["LDG"], c["STG"]))
        for k in keys:
            tot[k.split(".")[0]] += c[k]
    print("\n# totals: " + ", ".join("%s x%d" % kv for kv in sorted(tot.items())))


if __name__ == "__main__":
    sys.exit(main())
