#!/bin/bash
# round-2 (second session) check 4: parity + timings with the pipelined transient phase of the analysis kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cr_native.py -m gpu -x -q > gpurun_out/r02b_tests4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_tests4.log
tail -4 gpurun_out/r02b_tests4.log
( bash tools/quick.sh ) 2>&1 | grep pairs | tee gpurun_out/r02b_quick4.log
