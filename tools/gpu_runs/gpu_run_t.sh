#!/bin/bash
# round-2 GPU run T (2 GPUs): Shan-Chen psi planes stored straight into the neighbour's ghost planes -- IPC bit-identity, slab tests, self ring
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555"
timeout 300 $TR tools/slab_check.py > gpurun_out/r2t_slab_check_peer.txt 2>&1; echo "slab_check peer rc=$?"; tail -4 gpurun_out/r2t_slab_check_peer.txt
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_zv_peer_ring.py tests/test_gpu_zy_diag.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r2t_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log
tail -4 gpurun_out/r2t_pytest.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python tools/self_ring_bench.py sc3d 64 100 2>&1 | grep -v Warning > gpurun_out/r2t_self_ring.txt
cat gpurun_out/r2t_self_ring.txt
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2t_bench_n2.json 2> gpurun_out/r2t_bench_n2.err; echo "bench n2 rc=$?"
echo done
