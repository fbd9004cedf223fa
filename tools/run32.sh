#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1000 python -m pytest tests -m gpu -q 2>&1 | tail -12 ) > gpurun_out/r2_run32_pytest.txt
( timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_target.py 2>&1 | tail -12; echo "memcheck rc=$?" ) > gpurun_out/r2b_sanitizer_memcheck.txt
cat gpurun_out/r2_run32_pytest.txt gpurun_out/r2b_sanitizer_memcheck.txt
