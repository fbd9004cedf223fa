#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r2_run20_pytest.txt
o=gpurun_out/r2_run20_bucket.txt; : > $o
run() { echo "== $1 warps=$2 lo=$3" >> $o; DCRF_BUCKET_WARPS=$2 DCRF_BUCKET_LO=$3 timeout 300 python bench.py --config $1 --no-configs --no-sweep --no-cpu --steps 5 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
p = d['build_phases_ms_per_step']
print({k: d.get(k) for k in ('value', 'ms_per_step', 'build_ms_per_step')}, 'sort_d5', p.get('sort_d5'), 'csr_d5', p.get('csr_d5'), 'sort_d2', p.get('sort_d2'), 'csr_d2', p.get('csr_d2'))
" >> $o 2>&1; }
run voc32 8 8; run voc32 4 8; run voc32 2 8; run voc32 8 9; run voc32 4 9; run voc32 8 10
run dg2448 8 8; run dg2448 4 8; run adp1088_func 8 8; run adp1088_func 4 8; run hsn321x16 8 8; run hsn321x16 4 8
cat gpurun_out/r2_run20_pytest.txt $o
