#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r2_run18_pytest.txt
for cfg in voc32 dg2448 adp1088_func hsn321x16; do
timeout 300 python bench.py --config $cfg --no-configs --no-sweep --no-cpu --steps 5 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$cfg', {k: d.get(k) for k in ('value', 'ms_per_step', 'build_ms_per_step')}, d['build_phases_ms_per_step'])
"
done > gpurun_out/r2_run18_build.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"hash_insert_kernel<5|first_mask_kernel<5|assign_kernel<5|neighbour_wide_kernel<5|compact_insert_wide|bucket_csr|seg_radix|lattice_point_kernel<5|pack_fast|norm_s|ln_to_pm|pm_to_ln|softmax_unary" -c 24 -o gpurun_out/r2_voc32_build -f python tools/ncu_config.py voc32 1 > gpurun_out/r2_ncu_build.log 2>&1
cat gpurun_out/r2_run18_pytest.txt gpurun_out/r2_run18_build.txt; tail -3 gpurun_out/r2_ncu_build.log
