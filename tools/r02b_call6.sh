#!/bin/bash
# round-2 (second session) check 6: rows per shared-memory tile of the analysis kernels (the CMDR_SHT_TL_A* switches of that experiment were
# removed again after the measurement: 192 rows -0.3 %, 256 rows slower; profiles/r02_legendre_session2.txt)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( bash tools/quick.sh
  bash tools/quick.sh CMDR_SHT_TL_A2=192 CMDR_SHT_TL_A0=192
  bash tools/quick.sh CMDR_SHT_TL_A2=256 CMDR_SHT_TL_A0=256 ) 2>&1 | grep pairs | tee gpurun_out/r02b_quick6.log
CMDR_SHT_TL_A2=192 CMDR_SHT_TL_A0=256 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size or golden or adjoint or subset" 2>&1 | tail -3
