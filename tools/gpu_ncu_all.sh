#!/bin/bash
# ncu --set full capture of one whole step (13 launches) after warm-up. Output gpurun_out/prof_$1.ncu-rep
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -s ${2:-39} -c ${3:-13} -f -o gpurun_out/prof_$1 \
    python bench.py --steps 1 --warmup 3 --pairs ${4:-128} --no-cpu-baseline > gpurun_out/ncu_$1.log 2>&1
tail -2 gpurun_out/ncu_$1.log | cut -c1-300
ls -la gpurun_out/
