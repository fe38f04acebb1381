#!/bin/bash
# round-2 GPU run AD: HCZ D2Q9 multi-step (opt-in) parity; ncu full capture of puls_fused at N = 1024
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zq_hcz2d_multistep.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2ad_pytest.log 2>&1
tail -8 gpurun_out/r2ad_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:puls_fused --launch-skip 5 -c 1 -f -o gpurun_out/r2ad_puls_fused_1024 \
    python bench.py --workload c5_pulsatile_1024 --steps 3 --warmup 4 --no-e2e --no-cpu --no-extras > gpurun_out/r2ad_ncu.log 2>&1
tail -2 gpurun_out/r2ad_ncu.log
echo done
