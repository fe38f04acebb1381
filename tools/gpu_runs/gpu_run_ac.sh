#!/bin/bash
# round-2 GPU run AC: HCZ D2Q9 multi-step cooperative launch (configs[1]) -- parity + timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zq_hcz2d_multistep.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/r2ac_pytest.log 2>&1
tail -15 gpurun_out/r2ac_pytest.log
timeout 300 python tools/hcz2d_multi.py 2000 > gpurun_out/r2ac_hcz2d_multi.txt 2>&1; cat gpurun_out/r2ac_hcz2d_multi.txt
echo done
