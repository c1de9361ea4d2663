#!/usr/bin/env python
"""Latency of one sticky update call for ONE video (CUDA-graph replay), three measurements."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
dev = torch.device("cuda:0")
print([round(bench.run_single_video(dev)["us_per_call"], 2) for _ in range(3)])
