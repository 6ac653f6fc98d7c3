#!/bin/bash
for c in 1 2 3; do
  python bench.py --no-cpu-baseline --steps 12 --e2e-contexts $c 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('e2e contexts $c', 'resident %.0f e2e %.0f' % (d['value'], d['e2e']['value']))"
done
python bench.py --no-cpu-baseline --steps 12 --e2e-contexts 2 --pairs 128 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('pairs 128 contexts 2', 'resident %.0f e2e %.0f' % (d['value'], d['e2e']['value']))"
