#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r2_run28_pytest.txt
for cfg in voc32 adp1088_func dg2448; do
timeout 300 python bench.py --config $cfg --no-configs --no-sweep --no-cpu --steps 5 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$cfg', {k: d.get(k) for k in ('value', 'ms_per_step', 'build_ms_per_step')}, d['build_phases_ms_per_step'])
"
done > gpurun_out/r2_run28_build.txt 2>&1
cat gpurun_out/r2_run28_pytest.txt gpurun_out/r2_run28_build.txt
