#!/bin/bash
# round-2 GPU run AG: Pulsatile fused step with TMA-staged inputs -- bit-exact tests, stage / tile variants at N = 1024
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pulsatile.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/r2ag_pytest_puls.log 2>&1
tail -5 gpurun_out/r2ag_pytest_puls.log
for v in 0 2 3 4 5; do
  echo "CLBM_PULS_TMA=$v: $(CLBM_PULS_TMA=$v timeout 300 python bench.py --workload c5_pulsatile_1024 --steps 50 --warmup 5 --no-extras --no-e2e --no-cpu 2>/dev/null | python tools/pick.py)"
done | tee gpurun_out/r2ag_puls_variants.txt
for x in 40 52 80 104; do
  echo "CLBM_PULS_TMA=2 xchunk $x: $(CLBM_PULS_XCHUNK=$x timeout 300 python bench.py --workload c5_pulsatile_1024 --steps 50 --warmup 5 --no-extras --no-e2e --no-cpu 2>/dev/null | python tools/pick.py)"
done | tee -a gpurun_out/r2ag_puls_variants.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2ag_puls_launches.csv python bench.py --workload c5_pulsatile_1024 --steps 3 --warmup 3 --no-extras --no-e2e --no-cpu > /dev/null 2>&1
echo done
