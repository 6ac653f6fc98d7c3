#!/bin/bash
# small-batch side numbers (quadtree budget experiment): outputs in gpurun_out/small_$1.log
tag=${1:-s1}
mkdir -p gpurun_out
for a in "2160 3840 8000 1 50" "2160 3840 8000 8 20" "2160 3840 8000 32 10" "376 1241 2000 1 200" "376 1241 2000 8 100" "376 1241 10000 1 100" "376 1241 10000 16 50"; do timeout 300 python tools/bench_shape.py $a 2>&1 | tail -1; done | tee gpurun_out/small_${tag}.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "detect_parity or sort or 4k" 2>&1 | tail -3
