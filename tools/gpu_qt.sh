#!/bin/bash
for t in 64 128 192 256; do
  NAV24_QT_THREADS=$t python bench.py --no-cpu-baseline --steps 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('qt threads $t', 'resident %.0f e2e %.0f' % (d['value'], d['e2e']['value']), d['stage_ms_per_step']['quadtree_order']/4)"
done
NAV24_QT_THREADS=128 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
