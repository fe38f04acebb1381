#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zs_sc2d_multistep.py tests/test_gpu_parity.py tests/test_gpu_zu_dropin.py tests/test_drivers.py -m gpu -q --timeout 300 -p no:cacheprovider -k "sc or laplace or contact or layered or multi" > gpurun_out/r2v_pytest.log 2>&1
tail -5 gpurun_out/r2v_pytest.log
timeout 300 python tools/small_lattice_multi.py 2000 > gpurun_out/r2v_small_multi.txt 2>&1
cat gpurun_out/r2v_small_multi.txt
echo done
