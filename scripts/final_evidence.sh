#!/bin/bash
# One gpurun call for the end-of-round evidence: tests, smoke, the bench line in its variants, the ncu launch list of
# the bench command at the headline batch, one `ncu --set full` capture per kernel (scripts/ncu_all.sh).
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/summary.txt
if [ -z "$SKIP_TESTS" ]; then
  timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > $OUT/test_gpu.log 2>&1; echo "tests rc=$?" >> $OUT/summary.txt
  tail -2 $OUT/test_gpu.log >> $OUT/summary.txt
  timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/summary.txt
fi
SECONDS=0
timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$? wall=${SECONDS}s" >> $OUT/summary.txt
B="timeout 600 python bench.py --no-e2e --no-cpu-baseline --no-gauss --no-configs"
$B --no-overlap > $OUT/bench_noov.json 2> $OUT/bench_noov.err
$B --no-bin-pool > $OUT/bench_nobin.json 2> $OUT/bench_nobin.err
SECONDS=0
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$? wall=${SECONDS}s" >> $OUT/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-gauss --no-configs --sustain-s 0 > $OUT/ncu_launch.log 2>&1
echo "launch list rc=$?" >> $OUT/summary.txt
bash scripts/ncu_all.sh
# condense on the box: the .ncu-rep files of all kernels together exceed what gpurun_out/ may carry back
DEST=$OUT/profiles bash scripts/save_profiles.sh ${TAG:-final} > $OUT/save_profiles.log 2>&1
rm -f $OUT/prof_*.ncu-rep $OUT/launches.csv
timeout 300 python scripts/dropin_latency.py > $OUT/dropin_latency.txt 2>&1
cat $OUT/summary.txt
