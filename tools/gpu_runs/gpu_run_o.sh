#!/bin/bash
# round-2 GPU run O: signal / wait fused into the pack / unpack kernels -- tests and self-ring timings of the three protocols
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zv_peer_ring.py tests/test_gpu_zt_hcz3d_sweep.py tests/test_gpu_slab.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -5 gpurun_out/r2o_pytest.log
for k in "sc3d 64" "sc3d 512" "hcz2d 256" "hcz2d 2048" "hcz3d 64"; do timeout 300 python tools/self_ring_bench.py $k 100 2>&1 | grep -v Warning; done > gpurun_out/r2o_self_ring.txt
cat gpurun_out/r2o_self_ring.txt
echo done
