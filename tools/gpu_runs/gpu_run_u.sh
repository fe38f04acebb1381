#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_zv_peer_ring.py tests/test_gpu_zt_hcz3d_sweep.py tests/test_gpu_zw_overlap_hcz2d.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2u_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log
tail -4 gpurun_out/r2u_pytest.log
for k in "sc3d 64" "hcz2d 256" "hcz3d 64"; do
  timeout 300 python tools/self_ring_bench.py $k 100 2>&1 | grep -v Warning
  CLBM_RING_FUSE=0 timeout 300 python tools/self_ring_bench.py $k 100 2>&1 | grep -v Warning | sed 's/self ring/NOFUSE ring/'
done > gpurun_out/r2u_self_ring.txt
cat gpurun_out/r2u_self_ring.txt
echo done
