#!/usr/bin/env python
"""Per-kernel SASS evidence for the built library: counts of the tensor-core / TMEM / TMA mnemonics and of the 128-bit
global accesses in every kernel of libinfltm.so (cuobjdump -sass), written to profiles/<tag>_sass_summary.txt.

    python scripts/sass_summary.py r2
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "infinite_video_b200", "libinfltm.so")
WATCH = [("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCBAR (tcgen05.commit)", r"\bUTCBAR"),
         ("LDTM (tcgen05.ld)", r"\bLDTM"), ("STTM (tcgen05.st)", r"\bSTTM"), ("UTMALDG (TMA load)", r"\bUTMALDG"),
         ("UTMASTG (TMA store)", r"\bUTMASTG"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("LDG.128", r"\bLDG\.E[^ ]*\.128"),
         ("STG.128", r"\bSTG\.E[^ ]*\.128"), ("LDS.128", r"\bLDS[^ ]*\.128"), ("STS.128", r"\bSTS[^ ]*\.128"),
         ("MUFU.EX2", r"\bMUFU\.EX2"), ("REDUX", r"\bREDUX"), ("SHFL", r"\bSHFL"), ("HMMA/mma.sync", r"\bHMMA|\bIMMA")]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    sass = subprocess.check_output(["cuobjdump", "-sass", LIB], text=True, errors="replace")
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    total = collections.Counter()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            order.append(cur)
            continue
        if cur is None or "/*" not in line:
            continue
        ins = line.split("/*")[1] if line.strip().startswith("/*0") else line
        if not re.search(r"/\*[0-9a-f]{4}\*/", line):
            continue
        total[cur] += 1
        for name, pat in WATCH:
            if re.search(pat, line):
                counts[cur][name] += 1
    names = subprocess.run(["cu++filt"] + order, capture_output=True, text=True).stdout.splitlines() if order else []
    out = os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt")
    with open(out, "w") as f:
        f.write("cuobjdump -sass infinite_video_b200/libinfltm.so (sm_100a): per-kernel instruction counts\n")
        f.write("UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM -> registers), UTMALDG = cp.async.bulk.tensor (TMA), "
                "UTCBAR = tcgen05.commit, SYNCS = mbarrier ops\n\n")
        for mangled, nice in zip(order, names or order):
            c = counts[mangled]
            nice = re.sub(r"\(.*", "", nice)
            f.write(f"{nice[:110]}\n    instructions {total[mangled]:6d}")
            for name, _ in WATCH:
                if c[name]:
                    f.write(f" | {name} {c[name]}")
            f.write("\n")
        f.write("\nlibrary totals: " + ", ".join(
            f"{name} {sum(counts[k][name] for k in order)}" for name, _ in WATCH) + "\n")
    print(open(out).read()[-900:])


if __name__ == "__main__":
    main()
