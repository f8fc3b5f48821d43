#!/usr/bin/env python
"""Per-kernel histogram of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (no GPU needed).

    python tools/sass_opcodes.py > profiles/sass_opcodes_r02.txt

UTCHMMA = tcgen05.mma (kind::f16), UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st (TMEM), UTCCP = tcgen05.cp,
UTMALDG / UTMASTG = cp.async.bulk.tensor (TMA load / store), UBLKCP = cp.async.bulk, SYNCS = mbarrier, DFMA/DADD/DMUL = FP64.
"""
from __future__ import annotations

import re
import subprocess
import sys
from collections import Counter, defaultdict
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "orcai_b200" / "liborcai_b200.so"
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "FFMA", "DFMA", "DADD", "DMUL", "MUFU",
         "LDS", "STS", "LDG", "STG", "SHFL"]


def main() -> int:
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    demangled = dict(zip(re.findall(r"Function : (\S+)", sass), names))
    per = defaultdict(Counter)
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            per[cur][op] += 1
    print(f"# {LIB.name}: SASS mnemonic counts per kernel (cuobjdump -sass), sm_100a; columns: {' '.join(WATCH)} | total instructions")
    tot = Counter()
    for fn in sorted(per, key=lambda f: -sum(per[f].values())):
        c = per[fn]
        name = re.sub(r"\(anonymous namespace\)::|orcai::", "", demangled.get(fn, fn))
        name = re.sub(r"\(.*", "", name)[:110]
        row = " ".join(f"{c.get(w, 0):6d}" for w in WATCH)
        print(f"{row} | {sum(c.values()):7d}  {name}")
        tot.update(c)
    print("# whole library: " + ", ".join(f"{w} {tot.get(w, 0)}" for w in WATCH))
    return 0


if __name__ == "__main__":
    sys.exit(main())
