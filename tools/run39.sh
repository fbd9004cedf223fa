#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 70 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lsd_sort" 2>&1 | tail -6
