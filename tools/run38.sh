#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 85 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_target.py 2>&1 | tail -8; echo "racecheck rc=${PIPESTATUS[0]}" ) > gpurun_out/r2b_sanitizer_racecheck.txt
cat gpurun_out/r2b_sanitizer_racecheck.txt
