#!/bin/bash
# round-end check of the tree as it is: GPU parity suite, smoke, one default bench line.  Outputs in gpurun_out/
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/final_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/final_smoke.log
timeout 300 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench.json
