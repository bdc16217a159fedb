#!/bin/bash
# round-2 GPU call 5: position-major fold + two-level padding: parity subset, timing, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02_tests5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests5.log
tail -3 gpurun_out/r02_tests5.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-cg --no-batch --no-conviqt --no-parity --e2e-steps 3"
run() {
  name=$1; shift
  env "$@" timeout 300 $B > gpurun_out/r02_x_$name.json 2> gpurun_out/r02_x_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/r02_x_%s.json' % sys.argv[1]))
    k = d['roofline']['kernels']
    print(sys.argv[1], 'ms/pair %.3f' % d['ms_per_step'], 'fft', k['other_stages_ms'], 'sum %.3f' % sum(k['other_stages_ms'].values()),
          'e2e pageable %.2f pinned %.2f' % (d['e2e']['pageable']['ms_per_step'], d['e2e']['pinned']['ms_per_step']))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
{
run default X=1
run nobelt CMDR_SHT_BELT_FUSED=0
run split8192 CMDR_SHT_SPLIT_MIN=8192
run copy8 CMDR_SHT_COPY_THREADS=8
for sz in "1024 2000" "512 1500"; do
  echo "== pair_small $sz";  python tools/pair_small.py $sz | tail -1
done
} 2>&1 | tee gpurun_out/r02_fold_variants.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 320 --csv --log-file gpurun_out/r02_launches5.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cg --no-batch --no-conviqt --no-parity --e2e-steps 1 > gpurun_out/r02_ncu_l5.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ring_pow2_kernel|ring_split_kernel' -s 8 -c 4 -o gpurun_out/r02_ring_a -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-cg --no-batch --no-conviqt --no-parity --e2e-steps 1 > gpurun_out/r02_ncu_r.log 2>&1; echo "ncu full rc=$?"
