#!/bin/bash
# One gpurun call: environment probe, parity tests (grouped so a poisoned CUDA context cannot hide later
# groups), smoke, short bench, ncu launch list + one full capture of the dominant kernel.
# usage: scripts/gpu_check.sh [quick|full]
MODE=${1:-full}
OUT=gpurun_out
mkdir -p $OUT
{
  nvidia-smi -L; nproc; free -g | head -2; ls /root/reference 2>&1 | head -2
  python -c "import torch;print(torch.__version__, torch.cuda.get_device_name(0), torch.cuda.mem_get_info())"
} > $OUT/env.txt 2>&1
python __graft_entry__.py > $OUT/build.log 2>&1
run() { name=$1; shift; timeout 900 python -m pytest -q -m gpu -p no:cacheprovider "$@" > $OUT/test_$name.log 2>&1; echo "$name rc=$?" >> $OUT/summary.txt; tail -3 $OUT/test_$name.log >> $OUT/summary.txt; }
: > $OUT/summary.txt
run gemm tests/test_gpu_kernels.py -k "gemm or project_kv"
run kernels tests/test_gpu_kernels.py -k "not gemm and not project_kv"
run e2e_rect tests/test_gpu_e2e.py -k "not gauss"
run e2e_gauss tests/test_gpu_e2e.py -k "gauss"
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/summary.txt
if [ "$MODE" = "full" ]; then
  timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/summary.txt
  B="timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-gauss"
  $B --no-overlap > $OUT/bench_noov.json 2> $OUT/bench_noov.err
  $B --precision tf32x3 --no-overlap > $OUT/bench_x3.json 2> $OUT/bench_x3.err
  timeout 600 python bench.py --steps 2 --warmup 3 --videos 8 --no-e2e --no-cpu-baseline --gauss-videos 8 --gauss-frames 8192 > $OUT/bench_gauss_raw.json 2> $OUT/bench_gauss_raw.err
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
      python bench.py --steps 1 --warmup 3 --videos 32 --no-e2e --no-cpu-baseline --no-gauss > $OUT/ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:pool_mean -s 8 -c 2 -f -o $OUT/prof_pool \
      python bench.py --steps 1 --warmup 3 --videos 32 --no-e2e --no-cpu-baseline --no-gauss > $OUT/ncu_pool.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -s 8 -c 1 -f -o $OUT/prof_gemm \
      python bench.py --steps 1 --warmup 3 --videos 32 --no-e2e --no-cpu-baseline --no-gauss > $OUT/ncu_gemm.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:cont_attn -s 8 -c 1 -f -o $OUT/prof_attn \
      python bench.py --steps 1 --warmup 3 --videos 32 --no-e2e --no-cpu-baseline --no-gauss > $OUT/ncu_attn.log 2>&1
fi
cat $OUT/summary.txt
[ -f $OUT/bench.json ] && cat $OUT/bench.json; true
