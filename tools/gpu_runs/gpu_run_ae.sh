#!/bin/bash
# round-2 GPU run AE: Pulsatile with the node mask two columns ahead (bit-exact tests + bench); register-pipelined Shan-Chen kernel
# with unconditional halo loads (parity + slab tests); second ncu capture of the D2Q9 TMA kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pulsatile.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2ae_pytest_puls.log 2>&1
tail -3 gpurun_out/r2ae_pytest_puls.log
timeout 600 python bench.py --workload c5_pulsatile_1024 --steps 50 --warmup 5 --no-extras > gpurun_out/r2ae_bench_puls.json 2> gpurun_out/r2ae_bench_puls.err
python tools/pick.py < gpurun_out/r2ae_bench_puls.json
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab.py tests/test_gpu_zzz_sc_mrt.py tests/test_gpu_zs_sc2d_multistep.py -m gpu -q --timeout 600 -p no:cacheprovider -k "sc" > gpurun_out/r2ae_pytest_sc.log 2>&1
tail -3 gpurun_out/r2ae_pytest_sc.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2ae_puls_launches.csv python bench.py --workload c5_pulsatile_1024 --steps 3 --warmup 3 --no-extras --no-e2e --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sc2d_tma --launch-skip 3 -c 1 -f -o gpurun_out/r2ae_sc2d_8192 \
    python bench.py --workload sc_d2q9_8192 --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/r2ae_ncu.log 2>&1
tail -2 gpurun_out/r2ae_ncu.log
echo done
