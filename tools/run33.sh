#!/bin/bash
# final evidence A: GPU tests + ncu full capture of the iteration kernels of the final build (traffic json)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1000 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/r2c_pytest.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blur_kernel|splat_coop|splat_fast|slice_softmax_fast|splat_short" -s 14 -c 14 -o gpurun_out/r2c_voc32_hot -f python tools/ncu_config.py voc32 1 > gpurun_out/r2c_ncu_full.log 2>&1
cat gpurun_out/r2c_pytest.txt; tail -n 2 gpurun_out/r2c_ncu_full.log; ls -la gpurun_out
