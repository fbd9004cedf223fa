#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for cfg in adp1088_func adp1088_morph dg2448 voc32; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_$cfg.csv python tools/ncu_config.py $cfg 2 > gpurun_out/r2_ncu_$cfg.log 2>&1
done
echo done
