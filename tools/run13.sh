#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python tools/coresident_timing.py > gpurun_out/r2_run13_coresident.txt 2>&1
cat gpurun_out/r2_run13_coresident.txt
