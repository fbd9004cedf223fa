#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for cfg in adp1088_func adp1088_morph; do echo "== $cfg"; DCRF_TRACE=1 timeout 300 python tools/ncu_config.py $cfg 25 2>&1; done > gpurun_out/r2_run7_trace.txt
( timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu ) > gpurun_out/r2_run7_bench.json 2> gpurun_out/r2_run7_bench.err
echo done
