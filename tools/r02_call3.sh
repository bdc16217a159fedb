#!/bin/bash
# round-2 GPU call 3: new ring-FFT kernels (radix-4 split chirp-z, whole-ring belt FFT): parity suite + timing of the variants
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_comm_map.py tests/test_gpu_cr_native.py -m gpu -x -q > gpurun_out/r02_tests3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests3.log
tail -5 gpurun_out/r02_tests3.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-cg --no-batch --no-conviqt --no-parity --e2e-steps 1"
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 $B > gpurun_out/r02_v_$name.json 2> gpurun_out/r02_v_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/r02_v_%s.json' % sys.argv[1]))
    k = d['roofline']['kernels']
    print(sys.argv[1], 'ms/pair %.3f' % d['ms_per_step'], 'fft', {a: b for a, b in k['other_stages_ms'].items()}, 'sum %.3f' % sum(k['other_stages_ms'].values()))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
{
run default X=1
run nosplit CMDR_SHT_RING_SPLIT=0
run nobelt CMDR_SHT_BELT_FUSED=0
run old CMDR_SHT_RING_SPLIT=0 CMDR_SHT_BELT_FUSED=0 CMDR_SHT_FFT_BLOCKED=0
run unblocked CMDR_SHT_FFT_BLOCKED=0
run split8192 CMDR_SHT_SPLIT_MIN=8192
run split4096 CMDR_SHT_SPLIT_MIN=4096
run splitnt256 CMDR_SHT_SPLIT_NT=256
for sz in "1024 2000" "512 1500"; do
  echo "== pair_small $sz default";  python tools/pair_small.py $sz | tail -1
  echo "== pair_small $sz old";  CMDR_SHT_RING_SPLIT=0 CMDR_SHT_BELT_FUSED=0 CMDR_SHT_FFT_BLOCKED=0 python tools/pair_small.py $sz | tail -1
  echo "== pair_small $sz split4096";  CMDR_SHT_SPLIT_MIN=4096 python tools/pair_small.py $sz | tail -1
done
} 2>&1 | tee gpurun_out/r02_fft_variants.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches3.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cg --no-batch --no-conviqt --no-parity --e2e-steps 1 > gpurun_out/r02_ncu_l3.log 2>&1; echo "ncu launches rc=$?"
