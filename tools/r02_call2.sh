#!/bin/bash
# round-2 GPU call 2: full GPU suite, host-copy microbenchmark, bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests2.log
tail -5 gpurun_out/r02_tests2.log
for t in 1 2 4 8 12 16; do
  CMDR_SHT_COPY_THREADS=$t python - <<'PY'
import os, sys
sys.path.insert(0, '.')
from commander_b200 import sharp
L = sharp.lib()
n = 1 << 30
print("copy threads", L.cmdr_sht_host_copy_threads(), "up(wc) %.1f  up(plain) %.1f  down %.1f GB/s" % (
    L.cmdr_sht_measure_host_copy(n, 0, 1, 3), L.cmdr_sht_measure_host_copy(n, 0, 0, 3), L.cmdr_sht_measure_host_copy(n, 1, 0, 3)))
PY
done > gpurun_out/r02_hostcopy.log 2>&1
cat gpurun_out/r02_hostcopy.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_bench2.json'))
print({k: d[k] for k in ('value', 'ms_per_step')}, d['e2e']['pageable'], d['e2e']['pinned'], d['cg'], d.get('parity', {}).get('rel_l2'))
PY
CMDR_SHT_STAGE_WC=0 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-cg --no-batch --no-conviqt --no-parity > gpurun_out/r02_bench2_nowc.json 2>&1
python -c "
import json
d = json.load(open('gpurun_out/r02_bench2_nowc.json')); print('no WC:', d['e2e']['pageable'])"
