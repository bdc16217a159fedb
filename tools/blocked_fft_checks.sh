#!/bin/bash
# First GPU call of the next round: everything needed to make the register-blocked fused chirp-z kernel the default
# (DESIGN.md section 3.2).  Run on one B200:  gpurun --timeout 600 -- 'bash tools/blocked_fft_checks.sh'
# 1. the complete GPU suite under the switch; 2. pair timings at the three sizes, default vs blocked.
mkdir -p gpurun_out
CMDR_SHT_FFT_BLOCKED=1 timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/blocked_suite.log 2>&1
tail -3 gpurun_out/blocked_suite.log
for b in 0 1; do
  export CMDR_SHT_FFT_BLOCKED=$b
  echo "CMDR_SHT_FFT_BLOCKED=$b"
  python tools/pair_small.py 512 1500 | tail -1
  python tools/pair_small.py 1024 2000 | tail -1
  timeout 90 bash tools/quick.sh
done 2>&1 | tee gpurun_out/blocked_timing.log
