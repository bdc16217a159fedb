#!/bin/bash
# round-2 (second session) check 10: scalar front phase with the ring-dependent switch margin: parity + timings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cr_native.py tests/test_gpu_conviqt.py -m gpu -x -q > gpurun_out/r02b_tests10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_tests10.log
tail -6 gpurun_out/r02b_tests10.log
( bash tools/quick.sh ) 2>&1 | grep pairs | tee gpurun_out/r02b_quick10.log
