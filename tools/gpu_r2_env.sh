#!/bin/bash
# A/B of one environment knob of the library: usage gpu_r2_env.sh NAME v1 v2 ...
mkdir -p gpurun_out
name=$1; shift
for v in "$@"; do
  env $name=$v python bench.py --no-cpu-baseline --no-copy-ceiling --steps 10 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$name=$v', 'value %.0f'%d['value'], {k:round(x,3) for k,x in d['stage_ms_per_step'].items()}, 'parity', d['parity_checked']['mismatches'])"
done | tee gpurun_out/env_$name.log
