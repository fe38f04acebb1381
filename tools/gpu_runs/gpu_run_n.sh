#!/bin/bash
# round-2 GPU run N: the three slab protocols on a self ring at strong- and weak-scaling slab sizes
mkdir -p gpurun_out
for k in "sc3d 64" "sc3d 128" "sc3d 512" "hcz2d 256" "hcz2d 2048" "hcz3d 64"; do timeout 300 python tools/self_ring_bench.py $k 100 2>&1 | grep -v Warning; done > gpurun_out/r2n_self_ring.txt
cat gpurun_out/r2n_self_ring.txt
echo done
