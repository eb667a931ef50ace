#!/bin/bash
# Final record of a round in ONE gpurun call: profiling pass (launch list + ncu --set full of the C2 step and of the C5
# kernels), the DRAM-traffic files the bench lines quote, then the parity suite, one bench line per config and the
# reference arm.  Usage: scripts/gpu_final.sh <tag>
set -u
TAG=${1:-r2c}
OUT=gpurun_out
mkdir -p $OUT
bash scripts/profile_round.sh $TAG
python scripts/summarize_ncu.py full $OUT/prof_$TAG.ncu-rep $OUT/ncu_full_${TAG}_summary.csv $OUT/roofline_traffic_r2.json && cp $OUT/roofline_traffic_r2.json profiles/roofline_traffic_r2.json
python scripts/summarize_ncu.py full $OUT/prof_c5_$TAG.ncu-rep $OUT/ncu_full_${TAG}_c5_summary.csv $OUT/roofline_traffic_r2_c5.json && cp $OUT/roofline_traffic_r2_c5.json profiles/roofline_traffic_r2_c5.json
python scripts/summarize_ncu.py launches $OUT/launches_$TAG.csv $OUT/launches_${TAG}_summary.txt "ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sweep --eager (round 2, final build; cold-cache, serialised: compare SHARES)"
bash scripts/gpu_round.sh $TAG tests nolaunch
