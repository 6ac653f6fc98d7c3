#!/bin/bash
# bench stage times for several values of one environment knob: $1 = variable, rest = values
var=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
    export $var=$v
    python bench.py --no-cpu-baseline --no-copy-ceiling --steps 10 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$var=$v', 'value %.0f'%d['value'], {k:round(x,3) for k,x in d['stage_ms_per_step'].items()}, 'parity', d['parity_checked']['mismatches'])"
done | tee gpurun_out/env_${var}.log
