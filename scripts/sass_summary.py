#!/usr/bin/env python
"""profiles/<tag>_sass_summary.txt: per-kernel counts of the SASS mnemonics that prove which hardware paths the
shipped librse.so uses (tcgen05 MMA = UTCHMMA, TMA tensor loads = UTMALDG, bulk copies = UBLKCP, TMEM loads =
LDTM, tcgen05 commit = UTCBAR, packed fp32 FMA = FFMA2, ...).  Runs here: cuobjdump needs no GPU.

    python scripts/sass_summary.py [tag]
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "rag_search_engine_b200" / "csrc" / "librse.so"
WATCH = ["UTCHMMA", "UTCQMMA", "UTCMMA", "UTMALDG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS",
         "FFMA2", "FFMA", "DFMA", "DMUL", "DADD", "HMNMX2", "FMNMX3", "FMNMX", "ATOMS", "ATOMG", "REDG", "RED", "LDG", "STG",
         "LDS", "STS", "SHFL", "BAR", "MUFU", "IMAD", "LOP3"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = re.sub(r"\(.*", "", cur)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur:
            op, mods = m.group(1), m.group(2)
            kernels[cur][op] += 1
            kernels[cur]["total"] += 1
            if op in ("UTCHMMA", "UTMALDG", "UTCBAR", "LDTM", "UBLKCP"):
                kernels[cur][op + mods] += 1
    out = [f"SASS summary of {LIB.relative_to(ROOT)} ({tag}); cuobjdump -sass, counts of STATIC instructions per kernel.",
           "tcgen05.mma = UTCHMMA (.2CTA = cta_group::2), TMA tensor load = UTMALDG, cp.async.bulk = UBLKCP, tcgen05.ld = LDTM,",
           "tcgen05.commit = UTCBAR, mbarrier = SYNCS, packed fp32 FMA = FFMA2.", ""]
    for name, c in kernels.items():
        picks = [f"{k}={c[k]}" for k in WATCH if c.get(k)]
        detail = [f"{k}={v}" for k, v in sorted(c.items()) if "." in k]
        out.append(f"{name}\n    total={c['total']}  " + " ".join(picks))
        if detail:
            out.append("    " + " ".join(detail))
    tot = collections.Counter()
    for c in kernels.values():
        tot.update({k: v for k, v in c.items() if "." not in k})
    out += ["", "whole library: " + " ".join(f"{k}={tot[k]}" for k in WATCH if tot.get(k)), f"kernels: {len(kernels)}"]
    dst = ROOT / "profiles" / f"{tag}_sass_summary.txt"
    dst.write_text("\n".join(out) + "\n")
    print(dst)


if __name__ == "__main__":
    main()
