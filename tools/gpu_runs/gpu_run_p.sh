#!/bin/bash
# round-2 GPU run P: whole GPU suite + default bench on the final code
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2p_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -6 gpurun_out/r2p_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2p_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1; echo "smoke rc=$?"
echo done
