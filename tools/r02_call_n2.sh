#!/bin/bash
# round-2 multi-GPU call (N = number of visible GPUs, 2 by default): distributed parity in both exchange modes, bench with
# parity/checksums, PCIe probe.  Usage: gpurun --gpus N -- 'bash tools/r02_call_n2.sh N'
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
{
for p2p in 1 0; do
  echo "== dist_check world=$N CMDR_SHT_P2P=$p2p"
  CMDR_SHT_P2P=$p2p timeout 900 $TR --nproc-per-node $N --master-port $((29500 + p2p)) tests/dist_check.py > gpurun_out/r02_dist_full_p2p$p2p.log 2>&1
  grep -E "DIST_CHECK|Error|error|assert|Traceback|File \"" gpurun_out/r02_dist_full_p2p$p2p.log | head -30
done
echo "== dist_check world=$N, NCCL barriers (CMDR_SHT_FLAG_BARRIER=0)"
CMDR_SHT_FLAG_BARRIER=0 timeout 900 $TR --nproc-per-node $N --master-port 29503 tests/dist_check.py 2>&1 | grep -E "DIST_CHECK|Error|error|assert" | tail -5
} > gpurun_out/r02_dist_n$N.log 2>&1
cat gpurun_out/r02_dist_n$N.log
timeout 900 $TR --nproc-per-node $N --master-port 29510 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_n$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
for ln in open('gpurun_out/r02_bench_n%s.json' % n):
    if ln.startswith('{'):
        d = json.loads(ln)
        print('N', d['n_gpus'], 'pairs/s %.2f' % d['value'], 'ms %.3f' % d['ms_per_step'], 'e2e', {k: round(d['e2e'][k]['value'], 2) for k in ('pageable', 'pinned')},
              'parity', d.get('parity', {}).get('rel_l2'), d.get('parity', {}).get('mode'), 'cg', d['cg'] and round(d['cg']['value'], 1))
        print(d['roofline']['kernels'])
        print({k: v['sum_sq'] for k, v in d['checksums'].items()})
PY
CMDR_SHT_FLAG_BARRIER=0 timeout 600 $TR --nproc-per-node $N --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cg --no-parity --e2e-steps 2 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('NCCL barriers: pairs/s %.2f ms %.3f' % (d['value'], d['ms_per_step']), d['roofline']['kernels']['other_stages_ms'])"
timeout 300 $TR --nproc-per-node $N --master-port 29512 tools/pcie_probe.py 512 2>&1 | grep PCIE_PROBE | tee gpurun_out/r02_pcie_n$N.log
