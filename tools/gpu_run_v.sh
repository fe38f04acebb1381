#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/sc3d_variants.py 512 20 > gpurun_out/r2v_sc3d_variants.txt 2>&1
CLBM_TMA_PROMO=2 timeout 600 python tools/sc3d_variants.py 512 20 24 29 2>&1 | sed 's/^variant/promo128 variant/' >> gpurun_out/r2v_sc3d_variants.txt
CLBM_TMA_PROMO=0 timeout 600 python tools/sc3d_variants.py 512 20 24 29 2>&1 | sed 's/^variant/promo0 variant/' >> gpurun_out/r2v_sc3d_variants.txt
cat gpurun_out/r2v_sc3d_variants.txt
echo done
