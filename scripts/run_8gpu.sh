#!/bin/bash
# 8-GPU weak-scaling bench (torchrun, one rank per GPU) + the reference arm launched the same way; run as: gpurun --gpus 8 -- bash scripts/run_8gpu.sh
mkdir -p gpurun_out; python __graft_entry__.py > gpurun_out/build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29621 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "weak8 rc=$?" > gpurun_out/summary.txt
timeout 600 $TR --master-port 29622 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/bench_ref_8gpu.json 2> gpurun_out/bench_ref_8gpu.err; echo "ref8 rc=$?" >> gpurun_out/summary.txt
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; lscpu | head -30 >> gpurun_out/topo.txt
cat gpurun_out/summary.txt; tail -2 gpurun_out/bench_8gpu.err
python - <<'PY'
import json
p=json.load(open('gpurun_out/bench_8gpu.json')); print('8gpu', round(p['value']), p['n_gpus'], p['roofline_step']['frac'], (p.get('value_sustained') or {}).get('value'), json.dumps(p.get('e2e'))[:700])
PY
