#!/bin/bash
# round-2 GPU run G: sweep kernel (3 barriers, no prefetch, base-pointer stores); slab diagnostics
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zt_hcz3d_sweep.py tests/test_gpu_zy_diag.py -m gpu -q --timeout 600 -p no:cacheprovider -s > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -6 gpurun_out/r2g_pytest.log
B="python bench.py --no-extras --no-cpu --no-e2e"
timeout 300 $B --workload c4_hcz_d3q19_512 --steps 20 --warmup 5 > gpurun_out/r2g_bench_hcz3d.json 2> gpurun_out/r2g_bench_hcz3d.err; echo "hcz3d rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2g_bench_hcz3d.json').read().strip().splitlines()[-1]); print('hcz3d', round(d['value']), d['ms_per_step'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hcz3d_sweep --launch-skip 2 -c 1 -f -o gpurun_out/r2g_hcz3d_sweep_512 \
    $B --workload c4_hcz_d3q19_512 --steps 2 --warmup 3 > gpurun_out/r2g_ncu.log 2>&1
ncu -i gpurun_out/r2g_hcz3d_sweep_512.ncu-rep --page details > gpurun_out/r2g_hcz3d_sweep_512_ncu_full.txt 2>&1
echo done
