#!/bin/bash
# ncu --set full of the thin contraction kernels inside bench.py (C3 vertex-major; pass c4 for the sample-major shapes)
for c in ${1:-c3}; do
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_thin' -s 2 -c 2 \
    -o gpurun_out/ncu_thin_$c -f python bench.py --config $c --no-cpu-baseline --no-sweep --eager --steps 1 --warmup 1 > gpurun_out/ncu_thin_$c.log 2>&1
tail -2 gpurun_out/ncu_thin_$c.log
done
