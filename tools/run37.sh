#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest_full.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest_full.txt
tail -4 gpurun_out/r2c_pytest_full.txt
