#!/bin/bash
# GPU parity suite, A/B of the variants given as arguments, then ncu --set full (with source) of the kernels matching $NCU_K
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_s2e.log
bash tools/gpu_r2_variants.sh "$@"
cp gpurun_out/variants.log gpurun_out/variants_s2e.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${NCU_K:-quadtree_kernel|match_window|blur_kernel}" -s ${NCU_S:-6} -c ${NCU_C:-3} -f -o gpurun_out/prof_s2e \
    python bench.py --steps 1 --warmup 3 --pairs 128 --no-cpu-baseline --no-copy-ceiling > gpurun_out/ncu_s2e.log 2>&1
tail -2 gpurun_out/ncu_s2e.log | cut -c1-200
