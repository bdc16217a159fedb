#!/bin/bash
# round-2 (second session) check 1: full GPU suite + kernel timings after the two-l-per-step spin-0 kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_tests1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_tests1.log
tail -15 gpurun_out/r02b_tests1.log
bash tools/quick.sh 2>&1 | tail -3 | tee gpurun_out/r02b_quick1.log
