#!/bin/bash
# ncu --set full capture of selected kernels (regex in $1), small batch. Output gpurun_out/prof_$2.ncu-rep
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${3:-3} -c ${4:-1} -f -o gpurun_out/prof_$2 \
    python bench.py --steps 1 --warmup 3 --pairs ${5:-128} --no-cpu-baseline > gpurun_out/ncu_$2.log 2>&1
tail -2 gpurun_out/ncu_$2.log | cut -c1-300
