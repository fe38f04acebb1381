#!/bin/bash
# round-2 GPU run AB: D2Q9 Shan-Chen TMA kernel with the node masks fetched one column ahead -- parity + bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zr_sc2d_tma.py tests/test_gpu_zz_sc_rt2d.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2ab_pytest.log 2>&1
tail -3 gpurun_out/r2ab_pytest.log
for w in sc_d2q9_8192 sc_rt2d_2048; do
timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-extras --no-e2e --no-cpu 2>/dev/null | tail -1 | python tools/pick.py
done
timeout 300 python tools/sc2d_variants.py > gpurun_out/r2ab_sc2d_variants.txt 2>&1; cat gpurun_out/r2ab_sc2d_variants.txt
echo done
