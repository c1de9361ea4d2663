#!/usr/bin/env python
"""Summarise `ncu --page source --csv` of one kernel: top SASS lines by stall samples + stall-reason totals.
usage: ncu -i X.ncu-rep --page source --csv | python scripts/ncu_stalls.py [top_n]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
body = []
for r in rows[hi + 1:]:
    if r and r[0] in ("Kernel Name", "Address"):
        break                                  # a second captured launch follows: summarise the first only
    if len(r) == len(h):
        body.append(r)
c = {n: i for i, n in enumerate(h)}
S = "Warp Stall Sampling (All Samples)"
tot = sum(float(r[c[S]] or 0) for r in body) or 1.0
reasons = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg = {n: sum(float(r[c[n]] or 0) for r in body) for n in reasons}
print("kernel:", rows[0][1] if rows[0] else "?", " samples:", int(tot))
print("stall reasons:", ", ".join(f"{n[6:]}={100 * v / tot:.1f}%" for n, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v / tot > 0.01))
for r in sorted(body, key=lambda r: -float(r[c[S]] or 0))[:top]:
    rs = sorted(((float(r[c[n]] or 0), n[6:]) for n in reasons), reverse=True)[:2]
    print(f"{100 * float(r[c[S]]) / tot:5.1f}%  x{r[c['Instructions Executed']]:>8}  {r[c['Source']].strip()[:70]:70s} {rs[0][1]}:{int(rs[0][0])} {rs[1][1]}:{int(rs[1][0])}")
