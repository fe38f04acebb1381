#!/bin/bash
# round-2 GPU run AH: full GPU suite + smoke + the driver's bench command + its ncu launch list + ncu full capture of the D2Q9 TMA kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 -p no:cacheprovider > gpurun_out/r2ah_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2ah_pytest.log
tail -6 gpurun_out/r2ah_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2ah_smoke.log 2>&1
tail -2 gpurun_out/r2ah_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2ah_bench.json 2> gpurun_out/r2ah_bench.err
echo "bench rc=$?"; python tools/pick.py < gpurun_out/r2ah_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2ah_bench_launches.csv python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu --no-e2e --no-extras > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sc2d_tma --launch-skip 3 -c 1 -f -o gpurun_out/r2ah_sc2d_8192 \
    python bench.py --workload sc_d2q9_8192 --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/r2ah_ncu.log 2>&1
tail -2 gpurun_out/r2ah_ncu.log
timeout 300 python bench.py --workload sc_d2q9_8192 --steps 30 --warmup 5 --no-extras > gpurun_out/r2ah_bench_sc2d.json 2>/dev/null; python tools/pick.py < gpurun_out/r2ah_bench_sc2d.json
timeout 300 python bench.py --workload c5_pulsatile_1024 --steps 50 --warmup 5 --no-extras > gpurun_out/r2ah_bench_puls.json 2>/dev/null; python tools/pick.py < gpurun_out/r2ah_bench_puls.json
echo done
