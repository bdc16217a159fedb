#!/bin/bash
# quick GPU check: parity subset, then bench kernel timings (optionally with env overrides given as args VAR=val)
for kv in "$@"; do export "$kv"; done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-cg --no-batch --no-conviqt --e2e-steps 1 2>&1 | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); k=j['roofline']['kernels']
print('$*', 'pairs/s %.2f ms %.2f e2e %.2f' % (j['value'], j['ms_per_step'], j['e2e']['ms_per_step']), {a:b for a,b in k.items() if a.endswith('_ms') or a.startswith('legendre')})"
