#!/bin/bash
set -x
mkdir -p gpurun_out
python scripts/cfg1_steps.py > gpurun_out/cfg1_plain.log 2>&1 || { tail gpurun_out/cfg1_plain.log; exit 1; }
cat gpurun_out/cfg1_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_cfg1.csv python scripts/cfg1_steps.py > gpurun_out/ncu_o.log 2>&1
python - <<'PY'
import csv
from collections import defaultdict
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_cfg1.csv')) if len(r)>5]
hdr=rows[0]; k=hdr.index('Kernel Name'); v=hdr.index('Metric Value')
agg=defaultdict(list)
for r in rows[1:]:
    try: agg[r[k][:60]].append(float(r[v].replace(',','')))
    except: pass
for n,vals in sorted(agg.items(), key=lambda x:-sum(x[1])): print(f'{n:60s} n={len(vals):4d} mean={sum(vals)/len(vals)/1e3:8.2f} us total={sum(vals)/1e6:8.3f} ms')
PY
