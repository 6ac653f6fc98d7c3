#!/bin/bash
# host pipeline variants: streams x taper x chunk
for cfg in "2 0 64" "2 1 64" "3 1 64" "4 1 64" "3 1 96" "3 1 48" "4 1 48"; do
  set -- $cfg
  NAV24_STREAMS=$1 NAV24_TAPER=$2 NAV24_CHUNK_FRAMES=$3 python bench.py --no-cpu-baseline --steps 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('streams $1 taper $2 chunk $3', 'resident %.0f e2e %.0f' % (d['value'], d['e2e']['value']))"
done
NAV24_TRACE=1 python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>&1 | grep trace | tail -2
