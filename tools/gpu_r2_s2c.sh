#!/bin/bash
# GPU parity suite on the working tree's library, then A/B of the prebuilt variants given as arguments
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_s2c.log
bash tools/gpu_r2_variants.sh "$@"
cp gpurun_out/variants.log gpurun_out/variants_s2c.log
