#!/usr/bin/env python
import json, sys, glob, os
for f in sorted(glob.glob(os.path.join(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out", "bench*.json"))):
    try:
        d = json.load(open(f))
        if d.get("impl") == "reference":
            print(os.path.basename(f), "reference", round(d["value"], 2)); continue
        print(f"{os.path.basename(f):18s} {d['value']:9.0f} chunks/s  step-roofline {d['roofline_step']['frac']:.3f}  ms/step {d['ms_per_step']:.3f}  "
              + " ".join(f"{k}={v:.3f}" for k, v in d["stage_ms_per_chunk_step"].items()) + f"  pool {d['roofline']['frac']:.2f}"
              + (f"  e2e {d['e2e']['value']:.0f}" if d.get("e2e") else ""))
    except Exception as e:
        print(os.path.basename(f), "ERR", e)
