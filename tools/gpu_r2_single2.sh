#!/bin/bash
# single-frame: stage times (device) and host-call latency of prebuilt variants
mkdir -p gpurun_out
for v in "$@"; do
    export NAV24_LIB=$PWD/variants/lib_$v.so
    for a in "376 1241 2000 1 300" "480 752 1000 1 300" "2160 3840 8000 1 100" "376 1241 2000 4 100"; do python tools/bench_shape.py $a 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['shape'], 'B', d['frames_per_step'], ['%.4f'%x for x in d['stage_ms_per_step']])"; done
    python tools/bench_latency.py 2>&1 | tail -2 | sed "s/^/$v /"
done | tee gpurun_out/single_variants2.log
