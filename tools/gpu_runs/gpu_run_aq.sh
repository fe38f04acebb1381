#!/bin/bash
# round-2 GPU run AQ: full GPU suite + smoke + the driver's bench command + ncu full capture of the HCZ D3Q19 sweep kernel with the TMA box stores
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 900 -p no:cacheprovider > gpurun_out/r2aq_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2aq_pytest.log
tail -4 gpurun_out/r2aq_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2aq_smoke.log 2>&1
tail -1 gpurun_out/r2aq_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2aq_bench.json 2> gpurun_out/r2aq_bench.err
echo "bench rc=$?"; python tools/pick.py < gpurun_out/r2aq_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hcz3d_sweep --launch-skip 3 -c 1 -f -o gpurun_out/r2aq_hcz3d_sweep_512 \
    python bench.py --workload c4_hcz_d3q19_512 --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/r2aq_ncu.log 2>&1
tail -2 gpurun_out/r2aq_ncu.log
echo done
