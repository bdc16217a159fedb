#!/bin/bash
# round-2 (second session) 8-GPU call: distributed parity at world 8 and bench at N = 8 (and N = 4) with the final kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29528 tests/dist_check.py > gpurun_out/r02b_dist_full_w8.log 2>&1
echo "world=8 rc=$?: $(grep -E 'DIST_CHECK' gpurun_out/r02b_dist_full_w8.log)" | tee gpurun_out/r02b_dist_n8.log; grep -E "Error|assert" gpurun_out/r02b_dist_full_w8.log | head -5
for N in 8 4; do
timeout 600 $TR --nproc-per-node $N --master-port $((29530 + N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02b_bench_n$N.json 2> gpurun_out/r02b_bench_n$N.err; echo "bench N=$N rc=$?"
python - $N <<'PY'
import json, sys
for ln in open('gpurun_out/r02b_bench_n%s.json' % sys.argv[1]):
    if ln.startswith('{'):
        d = json.loads(ln)
        print('N', d['n_gpus'], 'pairs/s %.2f' % d['value'], 'ms %.3f' % d['ms_per_step'], 'e2e', {k: round(d['e2e'][k]['value'], 2) for k in ('pageable', 'pinned')},
              'parity', d.get('parity', {}).get('rel_l2'), d.get('parity', {}).get('mode'), 'cg', d['cg'] and round(d['cg']['value'], 1))
        print({k: v for k, v in d['roofline']['kernels'].items() if not k.endswith('executed') and not k.endswith('nominal')})
        print({k: v['sum_sq'] for k, v in d['checksums'].items()})
PY
done
