#!/bin/bash
# A/B of prebuilt library variants (variants/lib_*.so, selected with NAV24_LIB): bench stage times + two small-batch shapes
mkdir -p gpurun_out
for v in "$@"; do
    export NAV24_LIB=$PWD/variants/lib_$v.so
    python bench.py --no-cpu-baseline --no-copy-ceiling --steps 10 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'value %.0f'%d['value'], {k:round(x,3) for k,x in d['stage_ms_per_step'].items()}, 'parity', d['parity_checked']['mismatches'])"
    for a in "376 1241 2000 1 200" "2160 3840 8000 1 50"; do python tools/bench_shape.py $a 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('   ', d['shape'], 'B', d['frames_per_step'], ['%.3f'%x for x in d['stage_ms_per_step']])"; done
done | tee gpurun_out/variants.log
