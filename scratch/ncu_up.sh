#!/bin/bash
# per-kernel durations of the up-sampler on a C2-like band: scratch/ncu_up.sh [dtype] [ratio]
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"upsample" -s 15 -c 9 --csv --log-file gpurun_out/up_$1_$2.csv python scratch/perf_up.py $1 $2 > /dev/null 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/up_$1_$2.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); mi=hdr.index('Metric Name')
d=collections.defaultdict(list)
for r in rows[1:]:
    d[(r[ki].split('(')[0][-40:], r[mi][:22])].append(float(r[vi].replace(',','')))
for k,v in d.items(): print('$1 $2', k, round(sum(v)/len(v)))
PY
