#!/bin/bash
# ncu --set full of the streaming recurrence step at C5 size (default form)
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_spmm' -s 40 -c 2 \
    -o gpurun_out/ncu_c5_${1:-tile} -f python bench.py --config c5 --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_c5_${1:-tile}.log 2>&1
tail -2 gpurun_out/ncu_c5_${1:-tile}.log
