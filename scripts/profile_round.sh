#!/bin/bash
# Profiling pass of a round (run on the GPU box through gpurun): ncu launch list of one C2 training step, one
# `--set full` capture of every native kernel of that step, one of the C5 streaming step.  Reports land in gpurun_out/;
# scripts/summarize_ncu.py turns them into the tracked summaries under profiles/ (and the roofline traffic file
# bench.py quotes).  Usage: scripts/profile_round.sh <tag>
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sweep --eager"
timeout 300 $CMD > $OUT/bench_eager_$TAG.json 2> $OUT/bench_eager_$TAG.err || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_l_$TAG.log 2>&1
echo "launch list exit $?"
# the native kernels of ONE step (the 5th of the process: after the eager count, warm-up and capture-free steps)
timeout 900 ncu --set full --clock-control none --import-source on \
    -k 'regex:k_(cheb_fused|cheb_clenshaw|dw_planes|dw_umma|dw_thin|contract_umma|basis_onchip|gemm_pipe|gemm_stream|pack_b|gemm_umma|bias_act_pool|softmax_xent|sgd_momentum)' -s 66 -c 22 \
    -o $OUT/prof_$TAG -f $CMD > $OUT/ncu_f_$TAG.log 2>&1
echo "full capture exit $?"; tail -1 $OUT/ncu_f_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_(spmm|gemm_stream)' -s 40 -c 4 \
    -o $OUT/prof_c5_$TAG -f python bench.py --config c5 --no-cpu-baseline --steps 1 --warmup 1 > $OUT/ncu_c5_$TAG.log 2>&1
echo "c5 capture exit $?"
ls -la $OUT/*.ncu-rep
