#!/bin/bash
# variants A/B, then resident chunking of the working tree's library on two streams
mkdir -p gpurun_out
bash tools/gpu_r2_variants.sh "$@"
for c in 0 512 342; do
  NAV24_RESIDENT_CHUNK=$c python bench.py --no-cpu-baseline --no-copy-ceiling --steps 10 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('resident chunk $c', 'resident %.0f e2e %.0f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done | tee gpurun_out/rchunk_s2f.log
