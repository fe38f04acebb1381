#!/bin/bash
# round-2 GPU run AF: D2Q9 Shan-Chen TMA kernel with the column box as ONE TMA dimension (up to 256 rows) instead of row pairs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zr_sc2d_tma.py tests/test_gpu_zz_sc_rt2d.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2af_pytest.log 2>&1
tail -3 gpurun_out/r2af_pytest.log
timeout 300 python tools/sc2d_variants.py > gpurun_out/r2af_sc2d_variants.txt 2>&1; cat gpurun_out/r2af_sc2d_variants.txt
echo done
