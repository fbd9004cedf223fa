#!/bin/bash
# evidence pass A: GPU tests, smoke, ncu full capture of the iteration kernels (traffic json) and of the
# build kernels, launch list of two steps, e2e chunk size check
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r2_run23_pytest.txt
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/r2_run23_smoke.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blur_kernel|splat_coop|splat_fast|slice_softmax_fast|splat_short" -s 14 -c 14 -o gpurun_out/r2b_voc32_hot -f python tools/ncu_config.py voc32 1 > gpurun_out/r2b_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lattice_point|hash_insert|first_mask|assign_kernel|neighbour_wide|compact_insert|bucket_csr|seg_radix|pack_fast|rep_kernel|scan_" -c 60 -o gpurun_out/r2b_voc32_build -f python tools/ncu_config.py voc32 1 > gpurun_out/r2b_ncu_build.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_voc32_launches_2_steps.csv python tools/ncu_config.py voc32 2 > /dev/null 2>&1
for c in 16 8; do
echo "== BENCH_CHUNK=$c"; BENCH_CHUNK=$c timeout 300 python bench.py --no-configs --no-sweep --no-cpu --steps 5 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step')}, 'e2e', d['e2e']['ms_per_step'], 'e2e_labels', d['e2e_labels']['ms_per_step'])
"; done > gpurun_out/r2_run23_chunk.txt 2>&1
cat gpurun_out/r2_run23_pytest.txt gpurun_out/r2_run23_smoke.txt gpurun_out/r2_run23_chunk.txt; tail -2 gpurun_out/r2b_ncu_full.log gpurun_out/r2b_ncu_build.log
