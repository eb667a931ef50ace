#!/bin/bash
# One GPU-box pass of a round: parity tests, one bench line per BASELINE config, the reference arm, launch list.
# Usage (through gpurun): scripts/gpu_round.sh <tag> [notests|tests] [nolaunch]
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
if [ "${2:-}" != "notests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_$TAG.log
  tail -3 $OUT/pytest_$TAG.log
fi
timeout 600 python bench.py > $OUT/bench_c2_$TAG.json 2> $OUT/bench_c2_$TAG.err; echo "bench c2 exit $?"
for c in c1 c3 c4 c5; do
  timeout 600 python bench.py --config $c > $OUT/bench_${c}_$TAG.json 2> $OUT/bench_${c}_$TAG.err; echo "bench $c exit $?"
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref_c2_$TAG.json 2> $OUT/bench_ref_c2_$TAG.err; echo "ref exit $?"
if [ "${3:-}" != "nolaunch" ]; then
  CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sweep --eager"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_l_$TAG.log 2>&1
  echo "ncu launches exit $?"
fi
head -c 600 $OUT/bench_c2_$TAG.json
