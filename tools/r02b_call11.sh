#!/bin/bash
# round-2 (second session) check 11: ring pairs of a thread 32 apart, slices skipped in the transient phase: parity + timings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cr_native.py tests/test_gpu_conviqt.py -m gpu -x -q > gpurun_out/r02b_tests11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_tests10.log
tail -6 gpurun_out/r02b_tests11.log
( bash tools/quick.sh ) 2>&1 | grep pairs | tee gpurun_out/r02b_quick11.log
