#!/bin/bash
# round-2 (second session) check 3: ring pairs per thread of the analysis kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( bash tools/quick.sh
  bash tools/quick.sh CMDR_SHT_R_A2=5 CMDR_SHT_MINB_A2=8 CMDR_SHT_R_A0=10 CMDR_SHT_MINB_A0=8
  bash tools/quick.sh CMDR_SHT_R_A2=6 CMDR_SHT_MINB_A2=8 CMDR_SHT_R_A0=12 CMDR_SHT_MINB_A0=8
  bash tools/quick.sh CMDR_SHT_R_A2=4 CMDR_SHT_MINB_A2=10 CMDR_SHT_R_A0=8 CMDR_SHT_MINB_A0=8 ) 2>&1 | grep pairs | tee gpurun_out/r02b_quick3.log
