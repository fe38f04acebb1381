#!/bin/bash
# round-2 GPU run A: whole GPU suite, default bench, ncu capture of the D2Q9 Shan-Chen kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sc_fused_kernel --launch-skip 3 -c 1 -f -o gpurun_out/r2a_sc2d_8192 \
    python bench.py --workload sc_d2q9_8192 --steps 3 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/r2a_ncu_sc2d.log 2>&1
ncu -i gpurun_out/r2a_sc2d_8192.ncu-rep --page details > gpurun_out/r2a_sc2d_8192_ncu_full.txt 2>&1
ncu -i gpurun_out/r2a_sc2d_8192.ncu-rep --page source --csv > gpurun_out/r2a_sc2d_8192_source.csv 2>&1
echo done
