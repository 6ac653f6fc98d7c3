#!/bin/bash
# per-kernel device times of one step (after warm-up): ncu gpu__time_duration only (cold-cache, serialised)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s ${1:-39} -c ${2:-13} --csv --log-file gpurun_out/times.csv \
    python bench.py --steps 1 --warmup 3 --pairs ${3:-128} --no-cpu-baseline > gpurun_out/times.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/times.csv')) if len(r)>5]
for i,r in enumerate(rows):
    if r[0]=='ID': hdr=r; rows=rows[i+1:]; break
ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
d=collections.OrderedDict()
for r in rows:
    k=(r[0],r[ki].split('(')[0][-22:]); v=float(r[vi].replace(',','')); u=r[ui]
    if u=='ns': v/=1e3
    if u=='ms': v*=1e3
    if u=='Mbyte': v*=1e6
    if u=='Kbyte': v*=1e3
    if u=='Gbyte': v*=1e9
    d.setdefault(k,{})[r[mi]]=v
for k,m in d.items():
    print(f"{k[1]:24s} {m.get('gpu__time_duration.sum',0):9.1f} us  inst {m.get('smsp__inst_executed.sum',0)/1e6:8.1f}M  dram rd {m.get('dram__bytes_read.sum',0)/1e6:7.1f} wr {m.get('dram__bytes_write.sum',0)/1e6:7.1f} MB")
PY
