#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2_run6_pytest.txt
( timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2_run6_bench.json 2> gpurun_out/r2_run6_bench.err
echo done
