#!/bin/bash
# round-2 8-GPU call: distributed parity at world 8 (and 4), bench at N = 8, PCIe / host ceiling with 8 ranks copying at once
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r02_topo_n8.log 2>&1
lscpu | grep -E "Model name|Socket|NUMA|Core|Thread" >> gpurun_out/r02_topo_n8.log
for W in 8 4; do
  timeout 600 $TR --nproc-per-node $W --master-port $((29520 + W)) tests/dist_check.py > gpurun_out/r02_dist_full_w$W.log 2>&1
  echo "world=$W rc=$?: $(grep -E 'DIST_CHECK' gpurun_out/r02_dist_full_w$W.log)"; grep -E "Error|assert" gpurun_out/r02_dist_full_w$W.log | head -5
done 2>&1 | tee gpurun_out/r02_dist_n8.log
timeout 600 $TR --nproc-per-node 8 --master-port 29530 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench rc=$?"
python - <<'PY'
import json
for ln in open('gpurun_out/r02_bench_n8.json'):
    if ln.startswith('{'):
        d = json.loads(ln)
        print('N', d['n_gpus'], 'pairs/s %.2f' % d['value'], 'ms %.3f' % d['ms_per_step'], 'e2e', {k: round(d['e2e'][k]['value'], 2) for k in ('pageable', 'pinned')},
              'parity', d.get('parity', {}).get('rel_l2'), d.get('parity', {}).get('mode'), 'cg', d['cg'] and round(d['cg']['value'], 1))
        print(d['roofline']['kernels'])
        print({k: v['sum_sq'] for k, v in d['checksums'].items()})
PY
timeout 300 $TR --nproc-per-node 8 --master-port 29531 tools/pcie_probe.py 256 2>&1 | grep PCIE_PROBE | tee gpurun_out/r02_pcie_n8.log
CMDR_SHT_FLAG_BARRIER=0 timeout 300 $TR --nproc-per-node 8 --master-port 29532 bench.py --gpus 8 --steps 20 --warmup 5 --no-cg --no-parity --e2e-steps 2 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('NCCL barriers: pairs/s %.2f ms %.3f' % (d['value'], d['ms_per_step']), d['roofline']['kernels']['other_stages_ms'])" | tee -a gpurun_out/r02_dist_n8.log
