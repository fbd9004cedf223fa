#!/bin/bash
# final multi-GPU line: N = $1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/r2c_bench_${N}gpu.json 2> gpurun_out/r2c_bench_${N}gpu.err
echo "rc=$?"; tail -c 400 gpurun_out/r2c_bench_${N}gpu.err; python -c "
import json
d = json.loads(open('gpurun_out/r2c_bench_${N}gpu.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('n_gpus', 'value', 'ms_per_step')}, d['e2e']['value'], d['e2e_labels']['value'], d['sweep']['images_per_s'], d['sweep']['confusion_sha256'][:16])
"
