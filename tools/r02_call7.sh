#!/bin/bash
# round-2 GPU call 7: pageable pipeline variants (download ring, chunk counts), aliasing test of the ring kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "aliasing or pinned or full_size_vs_oracle" > gpurun_out/r02_tests7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_tests7.log
tail -3 gpurun_out/r02_tests7.log
CMDR_SHT_DN_PIECE_MB=4 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size_vs_oracle or config1" > gpurun_out/r02_tests7dn.log 2>&1; echo "pytest(download ring) rc=$?"; tail -2 gpurun_out/r02_tests7dn.log
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-cg --no-batch --no-conviqt --no-parity --e2e-steps 3"
run() {
  name=$1; shift
  env "$@" timeout 300 $B > gpurun_out/r02_z_$name.json 2> gpurun_out/r02_z_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/r02_z_%s.json' % sys.argv[1]))
    print(sys.argv[1], 'ms/pair %.3f' % d['ms_per_step'], 'e2e pageable %.2f pinned %.2f' % (d['e2e']['pageable']['ms_per_step'], d['e2e']['pinned']['ms_per_step']))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
}
{
run default X=1
run dn4x6 CMDR_SHT_DN_PIECE_MB=4
run dn8x4 CMDR_SHT_DN_PIECE_MB=8 CMDR_SHT_DN_SLOTS=4
run dn16x3 CMDR_SHT_DN_PIECE_MB=16 CMDR_SHT_DN_SLOTS=3
run chunks16 CMDR_SHT_CHUNKS=16
run chunks16_dn4 CMDR_SHT_CHUNKS=16 CMDR_SHT_DN_PIECE_MB=4
run mch16 CMDR_SHT_MCHUNKS=16
} 2>&1 | tee gpurun_out/r02_pageable_variants.log
