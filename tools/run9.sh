#!/bin/bash
# sanitizers + ncu full capture of the hot kernels (voc32) + launch list + GPU tests + bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r2_run9_pytest.txt
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_target.py 2>&1 | tail -12; echo "memcheck rc=$?" ) > gpurun_out/r2_sanitizer_memcheck.txt
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_target.py 2>&1 | tail -12; echo "racecheck rc=$?" ) > gpurun_out/r2_sanitizer_racecheck.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_final_voc32_launches.csv python tools/ncu_config.py voc32 2 > gpurun_out/r2_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blur_kernel|splat_coop|splat_fast|slice_softmax_fast|splat_tail|splat_short" -s 15 -c 16 -o gpurun_out/r2_voc32_hot -f python tools/ncu_config.py voc32 1 > gpurun_out/r2_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blur_kernel|splat_coop|splat_fast|slice_softmax_ref|splat_tail|splat_short" -s 15 -c 16 -o gpurun_out/r2_voc32_hot_reference -f python tools/ncu_config.py voc32 1 reference > gpurun_out/r2_ncu_full_ref.log 2>&1
( timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2_run9_bench.json 2> gpurun_out/r2_run9_bench.err
echo done
