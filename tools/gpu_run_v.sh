#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/sc3d_variants.py 512 20 > gpurun_out/r2v_sc3d_variants.txt 2>&1
cat gpurun_out/r2v_sc3d_variants.txt
echo done
