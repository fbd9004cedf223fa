#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for cfg in adp1088_morph adp1088_func dg2448; do echo "== $cfg"; DCRF_TRACE=1 timeout 300 python tools/ncu_config.py $cfg 6 2>&1; done > gpurun_out/r2_run5_trace.txt
echo done
