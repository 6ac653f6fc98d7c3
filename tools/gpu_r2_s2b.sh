#!/bin/bash
# blur fork / staggered resident chunks: A/B on the resident number, then the GPU parity suite with the defaults
mkdir -p gpurun_out
run() { env "$@" python bench.py --no-cpu-baseline --no-copy-ceiling --steps 10 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$*', 'resident %.0f e2e %.0f ms %.3f parity %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['parity_checked']['mismatches']))"; }
{
run NAV24_BLUR_FORK=0 NAV24_STAGGER=0
run NAV24_BLUR_FORK=1 NAV24_STAGGER=0
run NAV24_BLUR_FORK=0 NAV24_STAGGER=1 NAV24_RESIDENT_CHUNK=512
run NAV24_BLUR_FORK=0 NAV24_STAGGER=1 NAV24_RESIDENT_CHUNK=256
run NAV24_BLUR_FORK=0 NAV24_STAGGER=1 NAV24_RESIDENT_CHUNK=128
run NAV24_BLUR_FORK=1 NAV24_STAGGER=1 NAV24_RESIDENT_CHUNK=512
run NAV24_BLUR_FORK=1 NAV24_STAGGER=1 NAV24_RESIDENT_CHUNK=256
run NAV24_BLUR_FORK=1 NAV24_STAGGER=1 NAV24_RESIDENT_CHUNK=128
run NAV24_BLUR_FORK=1 NAV24_STAGGER=1 NAV24_RESIDENT_CHUNK=256 NAV24_STREAMS=2
run NAV24_BLUR_FORK=1 NAV24_STAGGER=1 NAV24_RESIDENT_CHUNK=128 NAV24_STREAMS=2
} | tee gpurun_out/fork_s2b.log
python tools/bench_latency.py 2>&1 | tail -3 | tee gpurun_out/latency_s2b.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_s2b.log
