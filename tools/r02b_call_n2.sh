#!/bin/bash
# round-2 (second session) 2-GPU call: distributed parity + bench after the two-l-per-step spin-0 kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 2 --master-port 29551 tests/dist_check.py > gpurun_out/r02b_dist2_full.log 2>&1
echo "world=2 rc=$?: $(grep -E 'DIST_CHECK' gpurun_out/r02b_dist2_full.log)" | tee gpurun_out/r02b_dist_n2.log; grep -E "Error|assert" gpurun_out/r02b_dist2_full.log | head -5
timeout 600 $TR --nproc-per-node 2 --master-port 29552 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02b_bench_n2.json 2> gpurun_out/r02b_bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
for ln in open('gpurun_out/r02b_bench_n2.json'):
    if ln.startswith('{'):
        d = json.loads(ln)
        print('N', d['n_gpus'], 'pairs/s %.2f' % d['value'], 'ms %.3f' % d['ms_per_step'], 'e2e ms', {k: round(d['e2e'][k]['ms_per_step'], 2) for k in ('pageable', 'pinned')},
              'parity', d.get('parity', {}).get('rel_l2'), d.get('parity', {}).get('mode'), 'cg', d['cg'] and round(d['cg']['value'], 1))
        print(d['roofline']['kernels'])
        print({k: v['sum_sq'] for k, v in d['checksums'].items()})
PY
