#!/bin/bash
# smoke + bench (both arms) + ncu launch list; outputs in gpurun_out/
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -3 | tee gpurun_out/bench_ref.json
python bench.py 2>&1 | tail -5 | tee gpurun_out/bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
