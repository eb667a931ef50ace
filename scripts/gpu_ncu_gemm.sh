#!/bin/bash
# ncu --set full of the pipelined GEMM at the C5 contraction shape (inside bench.py --config c5)
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_gemm_pipe' -s 3 -c 3 \
    -o gpurun_out/ncu_gemm_c5 -f python bench.py --config c5 --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_gemm_c5.log 2>&1
tail -2 gpurun_out/ncu_gemm_c5.log
