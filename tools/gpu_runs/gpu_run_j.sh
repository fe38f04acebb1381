#!/bin/bash
# round-2 GPU run J (8 GPUs): the default bench invocation (weak headline + strong pass + bit-identity), and the round-1 transport for comparison
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2j_bench_n8.json 2> gpurun_out/r2j_bench_n8.err; echo "bench n8 rc=$?"; tail -3 gpurun_out/r2j_bench_n8.err
CLBM_SLAB_TRANSPORT=torch timeout 300 $TR bench.py --gpus 8 --steps 100 --warmup 5 --scaling strong --no-extras --no-e2e --no-cpu > gpurun_out/r2j_n8_strong_sc3d_torch.json 2> gpurun_out/r2j_n8_strong_sc3d_torch.err; echo "torch rc=$?"
CLBM_SLAB_TRANSPORT=torch timeout 300 $TR bench.py --gpus 8 --steps 200 --warmup 5 --scaling strong --workload c3_hcz_d2q9_full --no-extras --no-e2e --no-cpu > gpurun_out/r2j_n8_strong_c3_torch.json 2> gpurun_out/r2j_n8_strong_c3_torch.err; echo "torch c3 rc=$?"
CLBM_SLAB_GRAPH=0 timeout 300 $TR bench.py --gpus 8 --steps 200 --warmup 5 --scaling strong --workload c3_hcz_d2q9_full --no-extras --no-e2e --no-cpu > gpurun_out/r2j_n8_strong_c3_nograph.json 2> gpurun_out/r2j_n8_strong_c3_nograph.err; echo "nograph c3 rc=$?"
echo done
