#!/bin/bash
# round-2 GPU run F: sweep kernel with two barriers per plane; x-chunk chooser on strong-scaling slab sizes (single GPU, wrap mode)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zt_hcz3d_sweep.py -m gpu -q --timeout 600 -p no:cacheprovider -s > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -6 gpurun_out/r2f_pytest.log
B="python bench.py --no-extras --no-cpu --no-e2e"
timeout 300 $B --workload c4_hcz_d3q19_512 --steps 20 --warmup 5 > gpurun_out/r2f_bench_hcz3d.json 2> gpurun_out/r2f_bench_hcz3d.err; echo "hcz3d rc=$?"
# strong-scaling slab sizes of an 8-GPU run, on one GPU: chunk chooser (new default) against the old chunk lengths
for xc in 0 24 32 64; do
  CLBM_SC_XCHUNK=$xc timeout 120 $B --workload c4_sc_d3q19_512 --size 64x512x512 --steps 200 --warmup 10 > gpurun_out/r2f_sc3d_64_xc$xc.json 2>/dev/null
done
for xc in 0 12 24 32 48; do
  CLBM_HCZ2D_XCHUNK=$xc timeout 120 $B --workload c3_hcz_d2q9_slab --steps 500 --warmup 20 > gpurun_out/r2f_hcz2d_256_xc$xc.json 2>/dev/null
done
for f in gpurun_out/r2f_sc3d_64_xc*.json gpurun_out/r2f_hcz2d_256_xc*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value']), round(d['ms_per_step']*1000,1),'us')"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hcz3d_sweep --launch-skip 2 -c 1 -f -o gpurun_out/r2f_hcz3d_sweep_512 \
    $B --workload c4_hcz_d3q19_512 --steps 2 --warmup 3 > gpurun_out/r2f_ncu.log 2>&1
ncu -i gpurun_out/r2f_hcz3d_sweep_512.ncu-rep --page details > gpurun_out/r2f_hcz3d_sweep_512_ncu_full.txt 2>&1
echo done
