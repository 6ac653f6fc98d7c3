#!/bin/bash
# quick GPU check: parity tests; output to gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
