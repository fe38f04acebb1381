#!/bin/bash
# round-2 GPU run Y: D3Q19 / SC-RT MRT operators, default bench line with the C1 / C2 entries
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_zzzz_mrt19.py tests/test_gpu_zx_hcz_mrt.py tests/test_gpu_zzz_sc_mrt.py -m gpu -q --timeout 600 -p no:cacheprovider -s > gpurun_out/r2y_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2y_pytest.log
tail -12 gpurun_out/r2y_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2y_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2y_bench.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])
for k,v in d['roofline']['also'].items(): print(k, v.get('mlups'), v.get('frac'), v.get('us_per_step'), v.get('cpu_baseline_reference'))
PY
echo done
