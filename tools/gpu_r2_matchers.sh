#!/bin/bash
# ncu --set full of the matcher kernels (integer-pipe evidence BASELINE.json's metric asks for) + the batch sweep; outputs in gpurun_out/
tag=${1:-m1}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'match_window|bf_knn2|bf_merge' -s 2 -c 10 -f \
    -o gpurun_out/prof_match_${tag} python tools/run_matchers.py > gpurun_out/ncu_match_${tag}.log 2>&1
tail -2 gpurun_out/ncu_match_${tag}.log
timeout 900 python bench.py --batch-sweep --no-cpu-baseline 2> gpurun_out/bench_sweep_${tag}.err | tail -1 > gpurun_out/bench_sweep_${tag}.json
python -c "
import json; d=json.load(open('gpurun_out/bench_sweep_${tag}.json'))
for r in d['batch_sweep']: print(r)
"
