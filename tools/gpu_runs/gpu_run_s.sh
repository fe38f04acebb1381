#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --no-extras --no-cpu --no-e2e"
timeout 200 $B --workload c4_sc_d3q19_512 --steps 30 --warmup 5 > gpurun_out/r2s_sc3d.json 2>/dev/null
timeout 200 $B --workload c3_hcz_d2q9_full --steps 50 --warmup 5 > gpurun_out/r2s_hcz2d.json 2>/dev/null
timeout 200 $B --workload c4_sc_d3q19_512 --size 64x512x512 --steps 200 --warmup 10 > gpurun_out/r2s_sc3d_64.json 2>/dev/null
timeout 200 $B --workload c3_hcz_d2q9_slab --steps 500 --warmup 20 > gpurun_out/r2s_hcz2d_256.json 2>/dev/null
for f in gpurun_out/r2s_*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value']), round(d['ms_per_step']*1000,1),'us', round(d['roofline']['frac'],4))"; done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider -k "hcz_rt2d or sc_d3q19 or golden" > gpurun_out/r2s_pytest.log 2>&1; tail -3 gpurun_out/r2s_pytest.log
echo done
