#!/bin/bash
# session-2 opener: resident chunking on several streams (256 / 342 / 512 frames per chunk), then ncu --set full with source of
# fast_band / blur / describe / resize8 for line-level instruction counts
mkdir -p gpurun_out
for c in 0 512 342 256 128; do
  NAV24_RESIDENT_CHUNK=$c python bench.py --no-cpu-baseline --no-copy-ceiling --steps 10 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('resident chunk $c', 'resident %.0f e2e %.0f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done | tee gpurun_out/rchunk_s2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fast_band|blur_kernel|describe_kernel|resize8" -s 20 -c 10 -f -o gpurun_out/prof_s2a \
    python bench.py --steps 1 --warmup 3 --pairs 128 --no-cpu-baseline --no-copy-ceiling > gpurun_out/ncu_s2a.log 2>&1
tail -2 gpurun_out/ncu_s2a.log | cut -c1-200
