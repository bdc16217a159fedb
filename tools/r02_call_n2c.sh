#!/bin/bash
# round-2 2-GPU call c: m-chunked a_lm transfers in the distributed host pipeline: dist_check + bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 2 --master-port 29551 tests/dist_check.py > gpurun_out/r02_dist2c_full.log 2>&1
echo "world=2 rc=$?: $(grep -E 'DIST_CHECK' gpurun_out/r02_dist2c_full.log)"; grep -E "Error|assert" gpurun_out/r02_dist2c_full.log | head -5
timeout 600 $TR --nproc-per-node 2 --master-port 29552 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2c.json 2> gpurun_out/r02_bench_n2c.err; echo "bench rc=$?"
python - <<'PY'
import json
for ln in open('gpurun_out/r02_bench_n2c.json'):
    if ln.startswith('{'):
        d = json.loads(ln)
        print('N', d['n_gpus'], 'pairs/s %.2f' % d['value'], 'ms %.3f' % d['ms_per_step'], 'e2e ms', {k: round(d['e2e'][k]['ms_per_step'], 2) for k in ('pageable', 'pinned')},
              'parity', d.get('parity', {}).get('rel_l2'), 'cg', d['cg'] and round(d['cg']['value'], 1))
PY
