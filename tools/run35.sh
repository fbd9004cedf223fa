#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( MALLOC_CHECK_=3 timeout 600 python -X faulthandler bench.py --steps 3 --warmup 3 --no-cpu ) > gpurun_out/r2_run35_a.json 2> gpurun_out/r2_run35_a.err; echo "rc=$?" >> gpurun_out/r2_run35_a.err
tail -c 3000 gpurun_out/r2_run35_a.err
