#!/bin/bash
# round-2 GPU call 6: coalesced belt fold timing; pageable upload ring variants
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ring_fft_paths or full_size_vs_oracle or pinned or all_jobs" > gpurun_out/r02_tests6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests6.log
tail -3 gpurun_out/r02_tests6.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-cg --no-batch --no-conviqt --no-parity --e2e-steps 3"
run() {
  name=$1; shift
  env "$@" timeout 300 $B > gpurun_out/r02_y_$name.json 2> gpurun_out/r02_y_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/r02_y_%s.json' % sys.argv[1]))
    k = d['roofline']['kernels']
    print(sys.argv[1], 'ms/pair %.3f' % d['ms_per_step'], 'fft sum %.3f' % sum(k['other_stages_ms'].values()), k['other_stages_ms'],
          'e2e pageable %.2f pinned %.2f' % (d['e2e']['pageable']['ms_per_step'], d['e2e']['pinned']['ms_per_step']))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
{
run ring4x6 X=1
run noring CMDR_SHT_UP_PIECE_MB=0
run ring16x4 CMDR_SHT_UP_PIECE_MB=16 CMDR_SHT_UP_SLOTS=4
run ring32x3 CMDR_SHT_UP_PIECE_MB=32 CMDR_SHT_UP_SLOTS=3
run ring8x4_t8 CMDR_SHT_UP_PIECE_MB=8 CMDR_SHT_UP_SLOTS=4 CMDR_SHT_COPY_THREADS=8
run nobelt CMDR_SHT_BELT_FUSED=0 CMDR_SHT_UP_PIECE_MB=0
} 2>&1 | tee gpurun_out/r02_ring_variants.log
