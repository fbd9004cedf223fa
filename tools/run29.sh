#!/bin/bash
# final single-GPU lines of the round
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2c_bench_reference.json 2> gpurun_out/r2c_bench_reference.err
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/r2c_smoke.txt
tail -c 300 gpurun_out/r2c_bench.err; cat gpurun_out/r2c_smoke.txt; python -c "
import json
d = json.loads(open('gpurun_out/r2c_bench.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'build_ms_per_step')}, d['roofline']['traffic'], d['e2e']['ms_per_step'])
"
