#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pulsatile.py tests/test_drivers.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2v_pytest.log 2>&1
tail -3 gpurun_out/r2v_pytest.log
timeout 600 python bench.py --workload c5_pulsatile_1024 --steps 50 --warmup 5 --no-extras > gpurun_out/r2v_bench_puls.json 2> gpurun_out/r2v_bench_puls.err
tail -2 gpurun_out/r2v_bench_puls.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v_bench_puls.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('kernel_ms'), d['e2e']['value'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2v_puls_launches.csv python bench.py --workload c5_pulsatile_1024 --steps 3 --warmup 3 --no-extras --no-e2e --no-cpu > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2v_puls_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-8:]: print(r[4][:40], r[-1], r[-2])
PY
echo done
