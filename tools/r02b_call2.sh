#!/bin/bash
# round-2 (second session) check 2: parity suite + kernel timings after the analysis store / flush changes; R sweep of the new spin-0 kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cr_native.py -m gpu -x -q > gpurun_out/r02b_tests2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_tests2.log
tail -4 gpurun_out/r02b_tests2.log
( bash tools/quick.sh
  bash tools/quick.sh CMDR_SHT_R_S0=6 CMDR_SHT_MINB_S0=12 CMDR_SHT_R_A0=6 CMDR_SHT_MINB_A0=12
  bash tools/quick.sh CMDR_SHT_R_S0=8 CMDR_SHT_MINB_S0=12 CMDR_SHT_R_A0=8 CMDR_SHT_MINB_A0=12
  bash tools/quick.sh CMDR_SHT_R_S0=8 CMDR_SHT_MINB_S0=8 CMDR_SHT_R_A0=8 CMDR_SHT_MINB_A0=8
  bash tools/quick.sh CMDR_SHT_R_S0=4 CMDR_SHT_MINB_S0=12 CMDR_SHT_R_A0=4 CMDR_SHT_MINB_A0=12 ) 2>&1 | grep pairs | tee gpurun_out/r02b_quick2.log
