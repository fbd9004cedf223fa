#!/bin/bash
# final evidence: GPU tests, smoke, ncu full capture (traffic json), bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r2_run12_pytest.txt
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/r2_run12_smoke.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blur_kernel|splat_coop|splat_fast|slice_softmax_fast|splat_tail|splat_short" -s 15 -c 16 -o gpurun_out/r2_voc32_hot -f python tools/ncu_config.py voc32 1 > gpurun_out/r2_ncu_full.log 2>&1
( timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2_run12_bench.json 2> gpurun_out/r2_run12_bench.err
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2_run12_bench_reference.json 2> gpurun_out/r2_run12_bench_reference.err
echo done
