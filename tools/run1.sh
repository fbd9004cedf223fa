#!/bin/bash
# GPU run 1 of round 2: parity of the reference arithmetic, its cost against the FMA kernels, the
# TMA gather4 micro-benchmark, and the other BASELINE configs' timings.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_run1_smi.txt
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r2_run1_pytest.txt
( timeout 120 ./tools/micro/gather4 2>&1 ) > gpurun_out/r2_micro_gather4.txt
( for m in fma reference; do DCRF_ARITHMETIC=$m timeout 300 python tools/iter_timing.py 32; done
  for v in r4 r5 r8 s4; do DCRF_ARITHMETIC=reference DCRF_B200_LIB=$PWD/wsss_analysis_b200/csrc/tune/libdcrf_$v.so timeout 300 python tools/iter_timing.py 32; done
) > gpurun_out/r2_run1_arith.txt 2>&1
( DCRF_ARITHMETIC=reference timeout 600 python tools/config_timing.py; DCRF_ARITHMETIC=fma timeout 600 python tools/config_timing.py ) > gpurun_out/r2_run1_configs.txt 2>&1
( timeout 300 python tools/phase_timing.py ) > gpurun_out/r2_run1_phase.txt 2>&1
echo done
