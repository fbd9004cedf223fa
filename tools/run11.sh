#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r2_run11_pytest.txt
for thr in 2 4 6; do for cfg in adp1088_morph adp1088_func dg612x8 hsn321x16; do
  echo "== short-row threshold $thr $cfg"; DCRF_SPLAT_SHORT_ROWS=$thr timeout 300 python bench.py --config $cfg --no-configs --no-sweep --no-cpu --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.0f ms/step %.3f arith %s'%(d['value'],d['ms_per_step'],d['arithmetic']), [(k['kernel'][:16].replace('slice_softmax_ke','slice'),k['avg_us']) for k in d['roofline']['per_kernel']])"
done; done > gpurun_out/r2_run11_short.txt 2>&1
echo done
