#!/bin/bash
# A/B of the thin (short-reduction) contraction kernels: parity tests, then C3 / C4 / C1 with CG_THIN=1 and 0
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q -k "filter or lstm or glstm or gemm or newsgroups or c3 or sparse or model or first_layer or cgcnn" > $OUT/pytest_thin.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_thin.log
tail -4 $OUT/pytest_thin.log
run() {
  local tag=$1 c=$2; shift 2
  env "$@" timeout 400 python bench.py --config $c --no-cpu-baseline --no-sweep > $OUT/bench_${c}_$tag.json 2> $OUT/bench_${c}_$tag.err
  echo "bench $c $tag exit $?"
  python - <<PY
import json
try:
    d=json.load(open('$OUT/bench_${c}_$tag.json'))
    k=d.get('kernels_ms_per_step',{})
    print('  ms_per_step %.4f value %.1f'%(d['ms_per_step'], d['value']), {n:round(v['ms_per_step'],4) for n,v in k.items() if v['ms_per_step']>0.02})
except Exception as e:
    print('  no line', e)
PY
}
for c in c3 c4 c1; do
  run thin1 $c CG_THIN=1
  run thin0 $c CG_THIN=0
done
