#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/sc3d_variants.py 512 20 29 41 > gpurun_out/r2v_sc3d_variants.txt 2>&1
CLBM_SC_PERSIST_STATIC=1 timeout 300 python tools/sc3d_variants.py 512 20 41 2>&1 | sed 's/^variant/static variant/' >> gpurun_out/r2v_sc3d_variants.txt
cat gpurun_out/r2v_sc3d_variants.txt
echo done
