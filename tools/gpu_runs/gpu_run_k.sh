#!/bin/bash
# round-2 GPU run K: slab-step overhead on one GPU (self ring) at the strong-scaling slab sizes; slab tests with the fused copies
mkdir -p gpurun_out
for k in "sc3d 64" "hcz3d 64" "hcz2d 256"; do timeout 300 python tools/self_ring_bench.py $k 200 2>&1 | grep -v Warning; done > gpurun_out/r2k_self_ring.txt
cat gpurun_out/r2k_self_ring.txt
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_zv_peer_ring.py tests/test_gpu_zt_hcz3d_sweep.py tests/test_gpu_zw_overlap_hcz2d.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -5 gpurun_out/r2k_pytest.log
echo done
