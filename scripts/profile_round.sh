#!/bin/bash
# Profiling pass of a round (run on the GPU box through gpurun): plain bench first, then the ncu launch list and one
# full capture of every hot kernel of one training step.  Usage: scripts/profile_round.sh <tag>
set -u
TAG=${1:-r1_v4}
OUT=gpurun_out
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err || exit 1
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --eager"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_l_$TAG.log 2>&1
# one step's worth of the library's tensor-core / recurrence kernels, after two warm-up steps' worth
timeout 900 ncu --set full --clock-control none --import-source on \
    -k 'regex:k_(cheb_fused|cheb_clenshaw|dw_planes|dw_umma|dw_thin|contract_umma|basis_onchip|gemm_pipe)' -s 26 -c 13 \
    -o $OUT/prof_$TAG -f $CMD > $OUT/ncu_f_$TAG.log 2>&1
tail -2 $OUT/ncu_f_$TAG.log
