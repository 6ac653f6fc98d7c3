#!/bin/bash
# round-2 GPU iteration: parity tests, short bench, launch list; outputs in gpurun_out/ (tag = $1)
tag=${1:-iter}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_${tag}.log
timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_${tag}.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 45 -c 30 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 1 --warmup 3 --pairs 128 --no-cpu-baseline > gpurun_out/ncu_launch_${tag}.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches_${tag}.csv')) if len(r)>5]
hdr=None;acc=collections.OrderedDict()
for r in rows:
    if r[0]=='ID': hdr={h:i for i,h in enumerate(r)}; continue
    if hdr is None: continue
    try: v=float(r[hdr['Metric Value']].replace(',',''))
    except: continue
    k=r[hdr['Kernel Name']][:40]; acc.setdefault(k,[0,0]); acc[k][0]+=v; acc[k][1]+=1
for k,(v,n) in acc.items(): print(f'{k:42s} n={n:3d} total {v/1e3:9.1f} us')
PY
