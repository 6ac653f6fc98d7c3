#!/bin/bash
# round-2 final evidence run (tag = $1): parity tests, bench (both arms), seq64 smoke, launch list, ncu --set full of one step,
# single-frame latency, other shapes, batch sweep; outputs in gpurun_out/
tag=${1:-v3}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_${tag}.log
timeout 900 python bench.py 2> gpurun_out/bench_${tag}.err | tail -1 | tee gpurun_out/bench_${tag}.json | cut -c1-600
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref_${tag}.json | cut -c1-400
timeout 600 python bench.py --workload seq64 --sequences 8 --seq-frames 20 --steps 3 2>&1 | tail -1 | tee gpurun_out/bench_seq_${tag}.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 54 -c 36 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 1 --warmup 3 --pairs 128 --no-cpu-baseline --no-copy-ceiling > gpurun_out/ncu_launch_${tag}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -s 54 -c 18 -f -o gpurun_out/prof_${tag} \
    python bench.py --steps 1 --warmup 3 --pairs 128 --no-cpu-baseline --no-copy-ceiling > gpurun_out/ncu_${tag}.log 2>&1
python tools/bench_latency.py 2>&1 | tail -3 | tee gpurun_out/latency_${tag}.txt
for a in "480 752 1000 256 20" "480 640 1000 256 20" "2160 3840 8000 32 10" "376 1241 10000 128 10" "376 1241 2000 1 200" "2160 3840 8000 1 50"; do python tools/bench_shape.py $a 2>/dev/null | tail -1; done | tee gpurun_out/other_shapes_${tag}.jsonl | cut -c1-300
timeout 600 python bench.py --batch-sweep 2>/dev/null | tail -1 | tee gpurun_out/bench_batch_sweep_${tag}.json | cut -c1-400
tail -3 gpurun_out/bench_${tag}.err
ls -la gpurun_out/ | tail -12
