B="timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-gauss"
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do BENCH_DEBUG=1 $B > gpurun_out/bench_t.json 2>gpurun_out/bench_t.err; python -c "
import json
d=json.load(open('gpurun_out/bench_t.json')); c=d['clocks']; print($i, round(d['value']), round(d['host_enqueue_ms_per_step_eager'],2), c.get('max_query_us'))
"; grep "overlap host" gpurun_out/bench_t.err | tail -1; done
