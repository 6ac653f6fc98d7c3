#!/bin/bash
# parity tests + single-frame latency + 4K / other-shape side numbers; outputs in gpurun_out/ (tag = $1)
tag=${1:-x1}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_${tag}.log
timeout 300 python tools/bench_latency.py 2>&1 | tail -4 | tee gpurun_out/latency_${tag}.log
for a in "480 752 1000 512" "480 640 1000 512" "2160 3840 8000 32" "376 1241 10000 256"; do timeout 300 python tools/bench_shape.py $a 2>&1 | tail -1; done | tee gpurun_out/shape_${tag}.log
