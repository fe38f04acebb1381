#!/bin/bash
# round-2 GPU run AA: ncu full capture + launch list of the D2Q9 Shan-Chen TMA kernel at 8192^2, bench line of that workload
mkdir -p gpurun_out
timeout 300 python bench.py --workload sc_d2q9_8192 --steps 30 --warmup 5 --no-extras > gpurun_out/r2aa_bench_sc2d.json 2> gpurun_out/r2aa_bench_sc2d.err
tail -c 600 gpurun_out/r2aa_bench_sc2d.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2aa_sc2d_launches.csv python bench.py --workload sc_d2q9_8192 --steps 3 --warmup 3 --no-extras --no-e2e --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sc2d_tma --launch-skip 3 -c 1 -f -o gpurun_out/r2aa_sc2d_8192 \
    python bench.py --workload sc_d2q9_8192 --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/r2aa_ncu.log 2>&1
tail -2 gpurun_out/r2aa_ncu.log
echo done
