#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r2_run22_pytest.txt
o=gpurun_out/r2_run22_bench.txt; : > $o
for cfg in voc32 voc1 sec41x32 hsn321x16 adp1088_morph; do
timeout 300 python bench.py --config $cfg --no-configs --no-sweep --no-cpu --steps 10 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$cfg', {k: d.get(k) for k in ('value', 'ms_per_step', 'build_ms_per_step', 'gpu_launches')}, [(k['kernel'][:14], k['avg_us']) for k in d['roofline']['per_kernel']])
" >> $o 2>&1
done
cat gpurun_out/r2_run22_pytest.txt $o
