#!/bin/bash
# round-2 GPU run C: sweep kernel with deferred edge merges, slab-mode sweep, D2Q9 L2 prefetch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zt_hcz3d_sweep.py tests/test_gpu_zv_peer_ring.py tests/test_gpu_slab.py -m gpu -q --timeout 600 -p no:cacheprovider -s > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -15 gpurun_out/r2e_pytest.log
timeout 300 python bench.py --workload c4_hcz_d3q19_512 --steps 20 --warmup 5 --no-extras --no-cpu > gpurun_out/r2e_bench_hcz3d.json 2> gpurun_out/r2e_bench_hcz3d.err
echo "bench hcz3d rc=$?"; tail -3 gpurun_out/r2e_bench_hcz3d.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hcz3d_sweep --launch-skip 2 -c 1 -f -o gpurun_out/r2e_hcz3d_sweep_512 \
    python bench.py --workload c4_hcz_d3q19_512 --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/r2e_ncu.log 2>&1
ncu -i gpurun_out/r2e_hcz3d_sweep_512.ncu-rep --page details > gpurun_out/r2e_hcz3d_sweep_512_ncu_full.txt 2>&1
echo done
