#!/bin/bash
# Copies the judged evidence of a gpurun call from gpurun_out/ (scratch) into profiles/ (tracked).
# usage: scripts/save_profiles.sh <tag>      e.g. r1c
TAG=${1:?tag}
DEST=${DEST:-profiles}        # DEST=gpurun_out/profiles on the GPU box (only gpurun_out/ travels back, capped at 64 MiB)
mkdir -p $DEST
for f in gpurun_out/bench*.json; do [ -s "$f" ] && cp "$f" $DEST/${TAG}_$(basename $f); done
[ -f gpurun_out/launches.csv ] && DEST=$DEST python - "$TAG" <<'PY'
import csv, collections, os, sys
tag = sys.argv[1]
dest = os.environ.get('DEST', 'profiles')
rows = list(csv.reader(open('gpurun_out/launches.csv', errors='replace')))
hdr = None; agg = collections.defaultdict(list); order = []
for r in rows:
    if len(r) > 5 and r[0] == 'ID': hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get('Metric Name') == 'gpu__time_duration.sum':
            agg[d['Kernel Name']].append(float(d['Metric Value'].replace(',', '')))
            order.append((d['ID'], d['Kernel Name'][:90], d['Metric Value'], d.get('Metric Unit', '')))
tot = sum(sum(v) for v in agg.values()) or 1
with open(f'{dest}/{tag}_launches_summary.txt', 'w') as f:
    f.write('ncu --metrics gpu__time_duration.sum --clock-control none -c 400  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-gauss --no-configs --sustain-s 0  (128 videos)\n')
    f.write('(cold-cache, serialised launches: compare SHARES, not absolutes; torch randn/fill kernels are input generation outside the timed region)\n\n')
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"{k[:100]:100s} n={len(v):4d} avg_us={sum(v)/len(v)/1e3:9.1f} share={100*sum(v)/tot:5.1f}%\n")
    f.write('\nlaunch list (first 120):\n')
    for o in order[:120]: f.write(' '.join(map(str, o)) + '\n')
PY
for rep in gpurun_out/prof_*.ncu-rep; do
  [ -f "$rep" ] || continue
  k=$(basename $rep .ncu-rep); k=${k#prof_}
  ncu -i $rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]
keep=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','dram__cycles_active.avg','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__shared_mem_per_block_dynamic','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','smsp__inst_executed.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for row in r[2:]:
    d=dict(zip(h,row))
    for w in keep:
        if w in d: print(f'{w:75s} {d[w]:>18s} {r[1][h.index(w)]}')
    print()
" > $DEST/${TAG}_ncu_$k.txt
  ncu -i $rep --page source --csv 2>/dev/null | python scripts/ncu_stalls.py 20 >> $DEST/${TAG}_ncu_$k.txt
done
ls -la $DEST | tail -20
