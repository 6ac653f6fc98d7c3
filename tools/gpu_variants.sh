#!/bin/bash
# bench every kernel variant in variants/ (plus the default build): prints stage times
for so in default variants/*.so; do
  if [ "$so" = default ]; then unset NAV24_LIB; else export NAV24_LIB=$PWD/$so; fi
  python bench.py --no-cpu-baseline --steps 10 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']
print('$so', 'fps %.0f e2e %.0f |' % (d['value'], d['e2e']['value']), ' '.join('%s=%.3f' % (k[:6], v) for k, v in s.items()))"
done
