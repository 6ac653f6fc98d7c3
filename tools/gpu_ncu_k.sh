#!/bin/bash
# ncu --set full of the kernels matching $1 in one step of the bench (256 frames); output gpurun_out/prof_$2.ncu-rep
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${3:-21} -c ${4:-7} -f -o gpurun_out/prof_$2 \
    python bench.py --steps 1 --warmup 3 --pairs 128 --no-cpu-baseline --no-copy-ceiling > gpurun_out/ncu_$2.log 2>&1
tail -2 gpurun_out/ncu_$2.log | cut -c1-200
