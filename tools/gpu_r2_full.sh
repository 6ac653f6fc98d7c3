#!/bin/bash
# round-2 GPU evidence run: parity tests, bench (both arms), seq64 smoke, ncu --set full of one step; outputs in gpurun_out/ (tag = $1)
tag=${1:-full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_${tag}.log
timeout 900 python bench.py 2> gpurun_out/bench_${tag}.err | tail -1 | tee gpurun_out/bench_${tag}.json | cut -c1-600
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref_${tag}.json | cut -c1-400
timeout 600 python bench.py --workload seq64 --sequences 8 --seq-frames 20 --steps 3 2>&1 | tail -1 | tee gpurun_out/bench_seq_${tag}.json | cut -c1-600
timeout 900 ncu --set full --clock-control none --import-source on -s 39 -c 13 -f -o gpurun_out/prof_${tag} \
    python bench.py --steps 1 --warmup 3 --pairs 128 --no-cpu-baseline --no-copy-ceiling > gpurun_out/ncu_${tag}.log 2>&1
tail -3 gpurun_out/bench_${tag}.err
ls -la gpurun_out/ | tail -8
