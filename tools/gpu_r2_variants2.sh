#!/bin/bash
mkdir -p gpurun_out
for v in "$@"; do
    export NAV24_LIB=$PWD/variants/lib_$v.so
    echo $v
    for a in "376 1241 2000 1 200" "376 1241 2000 8 100" "376 1241 2000 16 100" "376 1241 2000 24 100" "376 1241 2000 32 100" "2160 3840 8000 1 50" "2160 3840 8000 8 30" "2160 3840 8000 16 20" "2160 3840 8000 24 20" "2160 3840 8000 32 10"; do python tools/bench_shape.py $a 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('   ', d['shape'], 'B', d['frames_per_step'], 'fps %.0f'%d['frames_per_sec'], ['%.3f'%x for x in d['stage_ms_per_step']])"; done
done | tee gpurun_out/variants2.log
