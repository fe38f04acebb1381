#!/bin/bash
# round-2 GPU run M: overlap protocol v2 -- slab tests, self-ring timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_zv_peer_ring.py tests/test_gpu_zw_overlap_hcz2d.py tests/test_gpu_zx_hcz_mrt.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2m_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -5 gpurun_out/r2m_pytest.log
for k in "sc3d 64" "sc3d 512" "hcz2d 256" "hcz2d 2048"; do timeout 300 python tools/self_ring_bench.py $k 100 2>&1 | grep -v Warning; done > gpurun_out/r2m_self_ring.txt
cat gpurun_out/r2m_self_ring.txt
echo done
