#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
o=gpurun_out/r2_run30_blur_hint.txt; : > $o
for h in 0 1 2 3 0; do
echo "== DCRF_BLUR_HINT=$h" >> $o
DCRF_BLUR_HINT=$h timeout 300 python bench.py --no-configs --no-sweep --no-cpu --steps 5 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step')}, [(k['kernel'][:14], k['avg_us']) for k in d['roofline']['per_kernel']])
" >> $o 2>&1
done
cat $o
