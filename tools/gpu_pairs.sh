#!/bin/bash
for cfg in "128 3 64" "256 3 64" "512 3 64" "512 4 96" "256 4 48"; do
  set -- $cfg
  NAV24_STREAMS=$2 NAV24_CHUNK_FRAMES=$3 python bench.py --no-cpu-baseline --steps 10 --pairs $1 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('pairs $1 streams $2 chunk $3', 'resident %.0f e2e %.0f' % (d['value'], d['e2e']['value']))"
done
