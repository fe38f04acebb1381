#!/bin/bash
# round-2 GPU run R (8 GPUs): the default bench invocation on the final code, plus the other protocol forms for comparison
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2r_bench_n8.json 2> gpurun_out/r2r_bench_n8.err; echo "bench n8 rc=$?"; tail -3 gpurun_out/r2r_bench_n8.err
S="--gpus 8 --warmup 5 --scaling strong --no-extras --no-e2e --no-cpu"
CLBM_SLAB_OVERLAP=1 timeout 300 $TR bench.py $S --steps 100 > gpurun_out/r2r_n8_strong_sc3d_form1.json 2> gpurun_out/r2r_n8_strong_sc3d_form1.err; echo "form1 rc=$?"
CLBM_SLAB_OVERLAP=2 timeout 300 $TR bench.py $S --steps 100 > gpurun_out/r2r_n8_strong_sc3d_form2.json 2> gpurun_out/r2r_n8_strong_sc3d_form2.err; echo "form2 rc=$?"
CLBM_SLAB_OVERLAP=0 timeout 300 $TR bench.py $S --steps 200 --workload c3_hcz_d2q9_full > gpurun_out/r2r_n8_strong_c3_form0.json 2> gpurun_out/r2r_n8_strong_c3_form0.err; echo "c3 form0 rc=$?"
CLBM_SLAB_OVERLAP=1 timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 5 --no-extras --no-e2e --no-cpu > gpurun_out/r2r_n8_weak_sc3d_form1.json 2> gpurun_out/r2r_n8_weak_sc3d_form1.err; echo "weak form1 rc=$?"
echo done
