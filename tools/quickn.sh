#!/bin/bash
# multi-GPU bench summary: tools/quickn.sh N [VAR=val ...]
N=$1; shift
for kv in "$@"; do export "$kv"; done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --no-cg --e2e-steps 2 2>&1 | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read()); k=j['roofline']['kernels']
print('N=$N $*', 'pairs/s %.2f ms %.2f e2e %.2f' % (j['value'], j['ms_per_step'], j['e2e']['ms_per_step']), {a:b for a,b in k.items() if a.endswith('_ms') or a.startswith('legendre')})"
