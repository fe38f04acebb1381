#!/bin/bash
# round-2 GPU run D (2 GPUs): the peer-memory ring across processes (CUDA IPC), bit-identity, default bench with the strong pass
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/slab_check.py > gpurun_out/r2d_slab_check_peer.txt 2>&1; echo "slab_check peer rc=$?"; tail -4 gpurun_out/r2d_slab_check_peer.txt
timeout 300 $TR tools/slab_check.py --transport=nccl > gpurun_out/r2d_slab_check_nccl.txt 2>&1; echo "slab_check nccl rc=$?"; tail -4 gpurun_out/r2d_slab_check_nccl.txt
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2d_bench_n2.json 2> gpurun_out/r2d_bench_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/r2d_bench_n2.err
CLBM_SLAB_TRANSPORT=torch timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 5 --scaling strong --no-extras --no-e2e --no-cpu > gpurun_out/r2d_bench_n2_strong_torch.json 2> gpurun_out/r2d_bench_n2_strong_torch.err
CLBM_SLAB_GRAPH=0 timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 5 --scaling strong --no-extras --no-e2e --no-cpu > gpurun_out/r2d_bench_n2_strong_nograph.json 2> gpurun_out/r2d_bench_n2_strong_nograph.err
timeout 600 $TR bench.py --gpus 2 --steps 100 --warmup 5 --scaling strong --no-extras --no-e2e --no-cpu > gpurun_out/r2d_bench_n2_strong_peer.json 2> gpurun_out/r2d_bench_n2_strong_peer.err
echo done
