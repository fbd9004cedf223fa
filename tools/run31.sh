#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r2_run31_pytest.txt
o=gpurun_out/r2_run31_bench.txt; : > $o
for cb in 0 1; do for cfg in voc1 sec41x32; do
echo "== DCRF_CONCURRENT_BUILDS=$cb $cfg" >> $o
DCRF_CONCURRENT_BUILDS=$cb timeout 300 python bench.py --config $cfg --no-configs --no-sweep --no-cpu --steps 20 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'build_ms_per_step')}, 'e2e', d['e2e']['ms_per_step'], 'labels', d['e2e_labels']['ms_per_step'])
" >> $o 2>&1
done; done
python tools/small_host_timing.py voc1 >> $o 2>&1
cat gpurun_out/r2_run31_pytest.txt $o
