#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for cap in 24 48 96 256; do for cfg in hsn321x16 adp1088_morph voc32; do
  echo "== cap $cap $cfg"; DCRF_SPLAT_LONG_ROW=$cap DCRF_ARITHMETIC=fma timeout 300 python bench.py --config $cfg --no-configs --no-sweep --no-cpu --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.0f ms/step %.3f'%(d['value'],d['ms_per_step']), [(k['kernel'][:16],k['avg_us']) for k in d['roofline']['per_kernel']])"
done; done > gpurun_out/r2_run10_cap.txt 2>&1
echo done
