#!/usr/bin/env python
"""Count the SASS mnemonics that show which hardware paths each kernel uses:
    python scripts/sass_evidence.py psa_b200/libpsa_b200.so profiles/r01_sass_evidence.md
(UTCIMMA = tcgen05.mma kind::i8, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UBLKCP = bulk copy,
SYNCS = mbarrier, UCGABAR = cluster barrier, UTCBAR = tcgen05.commit, IDP = dp4a.)"""
import collections
import re
import subprocess
import sys

KEYS = ("UTCIMMA", "UTCHMMA", "UTCQMMA", "LDTM", "UTCBAR", "UTMALDG", "UBLKCP", "SYNCS", "UCGABAR", "HMMA", "IDP",
        "DFMA", "DADD", "DMUL", "F2F", "PRMT")


def main(lib, dst):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.defaultdict(collections.Counter), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"/\*[0-9a-f]{4,8}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
        if m and cur:
            for key in KEYS:
                if m.group(1).startswith(key):
                    counts[cur][key] += 1
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.split("\n")
    out = ["# SASS evidence (`cuobjdump -sass psa_b200/libpsa_b200.so`, sm_100a)", "",
           "Static instruction counts per kernel.  `UTCIMMA` = `tcgen05.mma.kind::i8`, `LDTM` = `tcgen05.ld`, `UTCBAR` =",
           "`tcgen05.commit`, `UTMALDG` = `cp.async.bulk.tensor` (TMA), `UBLKCP` = `cp.async.bulk`, `SYNCS` = mbarrier ops,",
           "`UCGABAR` = cluster barrier, `IDP` = `dp4a`; there is no `HMMA` (legacy `mma.sync`) anywhere.", "",
           "| kernel | " + " | ".join(KEYS) + " |", "|---|" + "---:|" * len(KEYS)]
    for mangled, name in sorted(zip(counts, names), key=lambda p: p[1]):
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        out.append(f"| `{name}` | " + " | ".join(str(counts[mangled].get(k, "")) for k in KEYS) + " |")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main(*sys.argv[1:3])
