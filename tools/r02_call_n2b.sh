#!/bin/bash
# round-2 2-GPU call b: dist_check in both exchange modes + NCCL barriers; chunk-count variants of the host pipeline; copy pool under a CPU limit
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{
for v in "CMDR_SHT_P2P=1" "CMDR_SHT_P2P=0" "CMDR_SHT_FLAG_BARRIER=0"; do
  env $v timeout 900 $TR --nproc-per-node 2 --master-port 29541 tests/dist_check.py > gpurun_out/r02_dist2_full.log 2>&1
  echo "world=2 $v rc=$?: $(grep -E 'DIST_CHECK' gpurun_out/r02_dist2_full.log)"; grep -E "Error|assert" gpurun_out/r02_dist2_full.log | head -5
done
} 2>&1 | tee gpurun_out/r02_dist_n2b.log
B="bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-cg --no-batch --no-conviqt --no-parity --e2e-steps 3"
for v in "CMDR_SHT_DIST_CHUNKS=4" "CMDR_SHT_DIST_CHUNKS=8" "CMDR_SHT_DIST_CHUNKS=16"; do
  env $v timeout 300 $TR --nproc-per-node 2 --master-port 29542 $B 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('$v', 'pairs/s %.2f' % d['value'], 'e2e ms pageable %.2f pinned %.2f' % (d['e2e']['pageable']['ms_per_step'], d['e2e']['pinned']['ms_per_step']))"
done 2>&1 | tee -a gpurun_out/r02_dist_n2b.log
echo "== 8 CPUs for both ranks (taskset -c 0-7): pool sized per rank" | tee -a gpurun_out/r02_dist_n2b.log
timeout 300 taskset -c 0-7 $TR --nproc-per-node 2 --master-port 29543 $B 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('taskset 0-7', 'pairs/s %.2f' % d['value'], 'e2e ms pageable %.2f pinned %.2f' % (d['e2e']['pageable']['ms_per_step'], d['e2e']['pinned']['ms_per_step']))" | tee -a gpurun_out/r02_dist_n2b.log
