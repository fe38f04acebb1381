#!/bin/bash
# round-2 GPU run W: split-group default of the D3Q19 Shan-Chen TMA kernel -- parity tests, slab tests, ncu full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab.py tests/test_gpu_zv_peer_ring.py -m gpu -q --timeout 600 -p no:cacheprovider -k "sc or d3q19 or sc3d or slab or ring" > gpurun_out/r2w_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2w_pytest.log
tail -6 gpurun_out/r2w_pytest.log
timeout 300 python tools/sc3d_variants.py 512 20 11 24 > gpurun_out/r2w_sc3d_variants.txt 2>&1
cat gpurun_out/r2w_sc3d_variants.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sc_fused_tma --launch-skip 3 -c 1 -f -o gpurun_out/r2w_sc3d_512 \
    python bench.py --workload c4_sc_d3q19_512 --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/r2w_ncu.log 2>&1
tail -2 gpurun_out/r2w_ncu.log
echo done
