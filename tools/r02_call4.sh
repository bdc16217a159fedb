#!/bin/bash
# round-2 GPU call 4: ring-major phase layout + m-major fold: parity + timing; DMMA microbenchmark; pageable pipeline trace
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mix.py tests/test_gpu_conviqt.py tests/test_gpu_cr_multi.py -m gpu -x -q > gpurun_out/r02_tests4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests4.log
tail -4 gpurun_out/r02_tests4.log
CMDR_SHT_PH_LAYOUT=m timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "all_jobs or ring_fft_paths or midsize or full_size_vs_oracle" > gpurun_out/r02_tests4m.log 2>&1; echo "pytest(m-major) rc=$?"; tail -2 gpurun_out/r02_tests4m.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-cg --no-batch --no-conviqt --no-parity --e2e-steps 2"
run() {
  name=$1; shift
  env "$@" timeout 300 $B > gpurun_out/r02_w_$name.json 2> gpurun_out/r02_w_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/r02_w_%s.json' % sys.argv[1]))
    k = d['roofline']['kernels']
    print(sys.argv[1], 'ms/pair %.3f' % d['ms_per_step'], {a: k[a] for a in k if a.endswith('_ms')}, 'fft', k['other_stages_ms'], 'sum %.3f' % sum(k['other_stages_ms'].values()),
          'e2e pageable %.2f pinned %.2f' % (d['e2e']['pageable']['ms_per_step'], d['e2e']['pinned']['ms_per_step']))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
{
run ringmajor X=1
run mmajor CMDR_SHT_PH_LAYOUT=m
run ringmajor_nobelt CMDR_SHT_BELT_FUSED=0
run ringmajor_split8192 CMDR_SHT_SPLIT_MIN=8192
for sz in "1024 2000" "512 1500"; do
  echo "== pair_small $sz ring-major";  python tools/pair_small.py $sz | tail -1
  echo "== pair_small $sz m-major";  CMDR_SHT_PH_LAYOUT=m python tools/pair_small.py $sz | tail -1
done
} 2>&1 | tee gpurun_out/r02_layout_variants.log
./tools/ubench/ubench5 2>&1 | tee gpurun_out/r02_ubench5.log
# pageable pipeline timeline
CMDR_SHT_PIPE_TRACE=1 python - > gpurun_out/r02_pipe_trace.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from commander_b200 import comm_map, comm_mapinfo
info = comm_mapinfo(None, 2048, 4000, 3, True)
h = comm_map(info)
h.alm[:] = np.random.default_rng(0).standard_normal(h.alm.shape)
for it in range(3):
    t0 = time.perf_counter(); h.Y(); t1 = time.perf_counter(); h.YtW(); t2 = time.perf_counter()
    print("pageable pair %d: Y %.2f ms, YtW %.2f ms" % (it, 1e3 * (t1 - t0), 1e3 * (t2 - t1)), file=sys.stderr)
PY
tail -14 gpurun_out/r02_pipe_trace.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 320 --csv --log-file gpurun_out/r02_launches4.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cg --no-batch --no-conviqt --no-parity --e2e-steps 1 > gpurun_out/r02_ncu_l4.log 2>&1; echo "ncu launches rc=$?"
