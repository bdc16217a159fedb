#!/bin/bash
# round-2 (second session) check 7: scalar front phase of the spin-2 kernels: parity + timings (A/B with CMDR_SHT_FRONT=0)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cr_native.py tests/test_gpu_conviqt.py -m gpu -x -q > gpurun_out/r02b_tests7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_tests7.log
tail -6 gpurun_out/r02b_tests7.log
( bash tools/quick.sh
  bash tools/quick.sh CMDR_SHT_FRONT=0
  bash tools/quick.sh CMDR_SHT_MINB_S2=12 ) 2>&1 | grep pairs | tee gpurun_out/r02b_quick7.log
