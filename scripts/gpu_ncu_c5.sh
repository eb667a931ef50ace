#!/bin/bash
# ncu --set full of the streaming recurrence step at C5 size, both step forms
for b in 0 1; do
  CG_SPMM_BLOCK=$b timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_spmm_step' -s 40 -c 2 \
    -o gpurun_out/ncu_c5_blk$b -f python bench.py --config c5 --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_c5_blk$b.log 2>&1
  tail -2 gpurun_out/ncu_c5_blk$b.log
done
ls -la gpurun_out/*.ncu-rep
