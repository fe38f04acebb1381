#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_zv_peer_ring.py tests/test_gpu_zw_overlap_hcz2d.py tests/test_gpu_zt_hcz3d_sweep.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2x_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log
tail -5 gpurun_out/r2x_pytest.log
for k in "sc3d 64" "hcz2d 256" "hcz3d 64"; do
  timeout 200 python tools/slab_profile.py $k 2>&1 | grep -v Warning
done > gpurun_out/r2x_slab_profile.txt
cat gpurun_out/r2x_slab_profile.txt
for k in "sc3d 64" "hcz3d 64" "hcz2d 256"; do timeout 300 python tools/self_ring_bench.py $k 100 2>&1 | grep -v Warning; done > gpurun_out/r2x_self_ring.txt
cat gpurun_out/r2x_self_ring.txt
echo done
