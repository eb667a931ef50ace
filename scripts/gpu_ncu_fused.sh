#!/bin/bash
# ncu --set full of the fused forward / Clenshaw / dW kernels at the C2 layer-2 shape (batch 1024), report in gpurun_out/
TAG=${1:-r2}
timeout 300 python scripts/prof_fused.py --iters 2 > gpurun_out/prof_fused_$TAG.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_(cheb_fused|cheb_clenshaw|dw_planes)' -s 3 -c 3 \
    -o gpurun_out/ncu_fused_$TAG -f python scripts/prof_fused.py --iters 2 > gpurun_out/ncu_fused_$TAG.log 2>&1
tail -3 gpurun_out/ncu_fused_$TAG.log; ls -la gpurun_out/ncu_fused_$TAG.ncu-rep
