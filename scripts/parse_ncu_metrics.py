import csv,sys
rows=list(csv.reader(open("gpurun_out/gemm_cycles.csv")))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
h=rows[hi]; c={n:i for i,n in enumerate(h)}
from collections import OrderedDict
d=OrderedDict()
for r in rows[hi+1:]:
    if len(r)!=len(h): continue
    d.setdefault(r[c["ID"]],{"k":r[c["Kernel Name"]][:34]})[r[c["Metric Name"]]]=r[c["Metric Value"]]
seen={}
for k,v in d.items():
    if "gemm" not in v["k"].lower(): continue
    f=lambda n: float(v[n].replace(",",""))
    t=f("gpu__time_duration.sum"); cy=f("sm__cycles_elapsed.max")
    print(k, v["k"], "us=%.1f MHz=%.0f"%(t/1e3, cy/t*1e3), " ".join("%s=%s"%(n.split(".")[0].replace("__","_")[-28:], v[n]) for n in v if n not in ("k","gpu__time_duration.sum","sm__cycles_elapsed.max")))
