#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
o=gpurun_out/r2_run21_small_host.txt; : > $o
python tools/small_host_timing.py voc1 >> $o 2>&1
python tools/small_host_timing.py sec41x32 >> $o 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_voc1_launches.csv python tools/ncu_config.py voc1 3 > /dev/null 2>&1
python tools/ncu_launch_summary.py gpurun_out/r2_voc1_launches.csv >> $o 2>&1
cat $o
