#!/bin/bash
# round-2 (second session) check 15: bench.py end to end after the last text changes (default flags of the driver, fewer steps)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02b_bench_check.json 2> gpurun_out/r02b_bench_check.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02b_bench_check.json'))
print({k: d[k] for k in ('metric', 'value', 'ms_per_step', 'gpu_launches', 'n_gpus', 'steps', 'warmup', 'dtype', 'scaling')}, 'e2e', d['e2e']['value'], 'parity', d['parity']['rel_l2'], 'cg', d['cg']['value'])
print(d['config']['legendre']); print(d['roofline']['executed_convention']); print(d['roofline']['traffic'], d['clocks'])
PY
tail -3 gpurun_out/r02b_bench_check.err
