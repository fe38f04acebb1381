#!/bin/bash
# round-2 GPU run AI: load-issue variants of the HCZ D3Q19 sweep kernel, its parity tests, the driver's bench command
mkdir -p gpurun_out
timeout 600 python tools/hcz3d_sweep_variants.py 512 20 0 1 2 3 > gpurun_out/r2ai_sweep_variants.txt 2>&1
cat gpurun_out/r2ai_sweep_variants.txt | tail -6
for v in 3; do
  CLBM_HCZ3D_SWEEP_VAR=$v timeout 600 python -m pytest tests/test_gpu_zt_hcz3d_sweep.py -m gpu -x -q --timeout 500 -p no:cacheprovider > gpurun_out/r2ai_pytest_sweep_var$v.log 2>&1
  echo "var $v pytest rc=$?"; tail -2 gpurun_out/r2ai_pytest_sweep_var$v.log
done
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2ai_bench.json 2> gpurun_out/r2ai_bench.err
echo "bench rc=$?"; python tools/pick.py < gpurun_out/r2ai_bench.json
echo done
