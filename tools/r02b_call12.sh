#!/bin/bash
# round-2 (second session) check 12: ring pairs per thread / warps per SM of the spin-2 kernels once more, with the final kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( bash tools/quick.sh CMDR_SHT_R_S2=3 CMDR_SHT_MINB_S2=12 CMDR_SHT_R_A2=3 CMDR_SHT_MINB_A2=12
  bash tools/quick.sh CMDR_SHT_R_S2=2 CMDR_SHT_MINB_S2=16 CMDR_SHT_R_A2=2 CMDR_SHT_MINB_A2=16
  bash tools/quick.sh CMDR_SHT_R_S2=4 CMDR_SHT_MINB_S2=12 CMDR_SHT_R_S0=2 CMDR_SHT_MINB_S0=16 ) 2>&1 | grep pairs | tee gpurun_out/r02b_quick12.log
