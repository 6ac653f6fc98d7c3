#!/bin/bash
# A/B of the prebuilt variants given as arguments, then ncu --set full (with source) of quadtree / match / describe of the working tree's library
mkdir -p gpurun_out
bash tools/gpu_r2_variants.sh "$@"
cp gpurun_out/variants.log gpurun_out/variants_s2d.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"quadtree_kernel|match_window|describe_kernel|order_kernel" -s 8 -c 4 -f -o gpurun_out/prof_s2d \
    python bench.py --steps 1 --warmup 3 --pairs 128 --no-cpu-baseline --no-copy-ceiling > gpurun_out/ncu_s2d.log 2>&1
tail -2 gpurun_out/ncu_s2d.log | cut -c1-200
