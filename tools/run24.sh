#!/bin/bash
# evidence pass A: ncu full capture of the iteration kernels (traffic json), launch list of two steps
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blur_kernel|splat_coop|splat_fast|slice_softmax_fast|splat_short" -s 14 -c 14 -o gpurun_out/r2b_voc32_hot -f python tools/ncu_config.py voc32 1 > gpurun_out/r2b_ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_voc32_launches_2_steps.csv python tools/ncu_config.py voc32 2 > /dev/null 2>&1
ls -la gpurun_out; tail -n 2 gpurun_out/r2b_ncu_full.log
