#!/bin/bash
# round-2 GPU run I: D2Q9 Shan-Chen kernel with three register sets (DEEP) against the round-1 form; whole GPU suite
mkdir -p gpurun_out
B="python bench.py --no-extras --no-cpu --no-e2e"
for v in 0 3; do
  CLBM_SC_TILE=$v timeout 120 $B --workload sc_d2q9_8192 --steps 30 --warmup 5 > gpurun_out/r2i_sc2d_v$v.json 2>/dev/null
  CLBM_SC_TILE=$v timeout 120 $B --workload sc_rt2d_2048 --steps 30 --warmup 5 > gpurun_out/r2i_scrt_v$v.json 2>/dev/null
done
for f in gpurun_out/r2i_sc*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value']), d['ms_per_step'], d['roofline']['frac'])"; done
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -8 gpurun_out/r2i_pytest.log
echo done
