#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( timeout 1000 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r2_run3_pytest.txt
for c in 1 2 3 4; do for cfg in voc1 sec41x32; do
  echo "== $cfg ctas/sm $c" ; DCRF_PERSISTENT_CTAS_PER_SM=$c timeout 300 python bench.py --config $cfg --no-configs --no-sweep --no-cpu --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.0f ms/step %.3f build %.3f e2e %.0f e2eL %.0f launches %d'%(d['value'],d['ms_per_step'],d['build_ms_per_step'],d['e2e']['value'],d['e2e_labels']['value'],d['gpu_launches'])); print(d['build_phases_ms_per_step'])"
done; done > gpurun_out/r2_run3_persistent.txt 2>&1
( echo "== persistent off"; for cfg in voc1 sec41x32; do DCRF_PERSISTENT_MAX_PIXELS=0 timeout 300 python bench.py --config $cfg --no-configs --no-sweep --no-cpu --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.0f ms/step %.3f build %.3f'%(d['value'],d['ms_per_step'],d['build_ms_per_step']))"; done ) >> gpurun_out/r2_run3_persistent.txt 2>&1
( timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu ) > gpurun_out/r2_run3_bench.json 2> gpurun_out/r2_run3_bench.err
echo done
