#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
o=gpurun_out/r2_run34_exit.txt; : > $o
for m in closed unclosed pending_closed pending_unclosed many pipeline; do
  python tools/exit_check.py $m >> $o 2>&1; echo "rc=$? ($m)" >> $o
done
DCRF_CONCURRENT_BUILDS=0 python tools/exit_check.py pipeline >> $o 2>&1; echo "rc=$? (pipeline, concurrent off)" >> $o
cat $o
