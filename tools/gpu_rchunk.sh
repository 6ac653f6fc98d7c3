#!/bin/bash
for c in 16 32 64 128; do
  NAV24_RESIDENT_CHUNK=$c python bench.py --no-cpu-baseline --steps 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('resident chunk $c', 'resident %.0f e2e %.0f' % (d['value'], d['e2e']['value']))"
done
