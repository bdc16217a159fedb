#!/bin/bash
# round-2 (second session) check 14: per-call breakdown of the host-buffer path (pinned) with the pipeline trace
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMDR_SHT_PIPE_TRACE=1 python tools/e2e_probe.py > gpurun_out/r02b_e2e_probe.log 2>&1
grep -v "cmdr_sht pipe" gpurun_out/r02b_e2e_probe.log | tail -12
grep "cmdr_sht pipe" gpurun_out/r02b_e2e_probe.log | tail -4 | cut -c1-1400
