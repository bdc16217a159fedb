#!/bin/bash
# round-2 (second session) 2-GPU call b: distributed parity with the final anal2_kernel (ring pairs 32 apart, slice skipping)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 2 --master-port 29551 tests/dist_check.py > gpurun_out/r02b_dist2b_full.log 2>&1
echo "world=2 rc=$?: $(grep -E 'DIST_CHECK' gpurun_out/r02b_dist2b_full.log)" | tee gpurun_out/r02b_dist_n2b.log; grep -E "Error|assert" gpurun_out/r02b_dist2b_full.log | head -5
timeout 300 $TR --nproc-per-node 2 --master-port 29552 bench.py --gpus 2 --steps 10 --warmup 3 --no-cg --no-batch --no-conviqt --e2e-steps 1 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('N 2 pairs/s %.2f ms %.3f parity' % (d['value'], d['ms_per_step']), d['parity']['rel_l2'], d['parity']['mode'], {k: v['sum_sq'] for k, v in d['checksums'].items()})" | tee -a gpurun_out/r02b_dist_n2b.log
