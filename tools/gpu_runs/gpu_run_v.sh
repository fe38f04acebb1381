#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zr_sc2d_tma.py tests/test_gpu_zz_sc_rt2d.py tests/test_gpu_parity.py tests/test_gpu_zu_dropin.py tests/test_gpu_zy_diag.py tests/test_gpu_zs_sc2d_multistep.py tests/test_gpu_zzz_sc_mrt.py -m gpu -q --timeout 600 -p no:cacheprovider -k "sc or laplace or contact or layered or rt or tma or diag" > gpurun_out/r2v_pytest.log 2>&1
tail -5 gpurun_out/r2v_pytest.log
for w in sc_d2q9_8192 sc_rt2d_2048; do
timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-extras --no-e2e --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['workload'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel'])"
done
echo done
