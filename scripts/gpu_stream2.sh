#!/bin/bash
# second A/B pass of the streaming GEMM: lower row threshold + equal-width column tiles; per-launch list of a C4 step
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "gemm or linear or lstm or glstm or stream or partition or c5 or fourier or contract" > $OUT/pytest_stream2.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_stream2.log
tail -4 $OUT/pytest_stream2.log
run() {  # tag, env..., config
  local tag=$1 c=$2; shift 2
  env "$@" timeout 400 python bench.py --config $c --no-cpu-baseline --no-sweep > $OUT/bench_${c}_$tag.json 2> $OUT/bench_${c}_$tag.err
  echo "bench $c $tag exit $?"
  python - <<PY
import json
try:
    d=json.load(open('$OUT/bench_${c}_$tag.json'))
    k=d.get('kernels_ms_per_step',{})
    print('  ms_per_step %.4f value %.1f'%(d['ms_per_step'], d['value']), {n:round(v['ms_per_step'],4) for n,v in k.items() if 'gemm' in n})
except Exception as e:
    print('  no line', e)
PY
}
for c in c4 c2 c5 c1; do
  run m129 $c CG_GEMM_STREAM=1
  run m512 $c CG_GEMM_STREAM=1 CG_GEMM_STREAM_MIN_M=512
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_c4_stream.csv python bench.py --config c4 --steps 2 --warmup 1 --no-cpu-baseline --no-sweep --eager > $OUT/ncu_l_c4.log 2>&1
echo "ncu exit $?"
