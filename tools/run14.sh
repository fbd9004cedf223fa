#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
o=gpurun_out/r2_run14_small.txt; : > $o
( timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r2_run14_pytest.txt
for cfg in voc1 sec41x32; do
  for g in 0 1000000; do
    echo "== $cfg DCRF_GRAPH_MAX_PIXELS=$g" >> $o
    DCRF_GRAPH_MAX_PIXELS=$g timeout 300 python bench.py --config $cfg --no-configs --no-sweep --no-cpu --steps 20 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'build_ms_per_step', 'gpu_launches')}, d['e2e']['ms_per_step'], d['e2e_labels']['ms_per_step'])
print([(k['kernel'], k['avg_us']) for k in d['roofline']['per_kernel']])
" >> $o 2>&1
  done
done
cat gpurun_out/r2_run14_pytest.txt $o
