#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r2_run19_pytest.txt
( timeout 600 python bench.py --steps 5 --warmup 3 --no-configs --no-cpu ) > gpurun_out/r2_run19_bench.json 2> gpurun_out/r2_run19_bench.err
python - <<'P'
import json
d = json.loads(open('gpurun_out/r2_run19_bench.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'build_ms_per_step')}, d['build_phases_ms_per_step'])
print('sweep', d.get('sweep'))
P
cat gpurun_out/r2_run19_pytest.txt; tail -3 gpurun_out/r2_run19_bench.err
