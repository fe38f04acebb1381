#!/bin/bash
mkdir -p gpurun_out
for k in "sc3d 64" "hcz2d 256"; do timeout 300 python tools/self_ring_bench.py $k 100 2>&1 | grep -v Warning; done > gpurun_out/r2l_self_ring.txt
cat gpurun_out/r2l_self_ring.txt
echo done
