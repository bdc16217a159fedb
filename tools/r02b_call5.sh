#!/bin/bash
# round-2 (second session) check 5: half-length belt kernel (ring_half_kernel): parity + timings against the whole-pair kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02b_tests5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_tests5.log
tail -6 gpurun_out/r02b_tests5.log
( bash tools/quick.sh
  bash tools/quick.sh CMDR_SHT_BELT_HALF=0 ) 2>&1 | grep pairs | tee gpurun_out/r02b_quick5.log
python tools/pair_small.py 1024 2000 2>&1 | tail -1
python tools/pair_small.py 512 1500 2>&1 | tail -1
