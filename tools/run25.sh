#!/bin/bash
# evidence pass B: full bench line (all configurations + sweep + cpu baseline), reference arm, ncu full of the build kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2b_bench_reference.json 2> gpurun_out/r2b_bench_reference.err
timeout 600 ncu --set full --clock-control none -k regex:"lattice_point|hash_insert|first_mask|assign_kernel|neighbour_wide|compact_insert|bucket_csr|seg_radix|pack_fast|rep_kernel" -s 14 -c 26 -o gpurun_out/r2b_voc32_build -f python tools/ncu_config.py voc32 1 > gpurun_out/r2b_ncu_build.log 2>&1
ls -la gpurun_out; tail -c 600 gpurun_out/r2b_bench.err; tail -n 2 gpurun_out/r2b_ncu_build.log
