#!/bin/bash
# A/B of the streaming GEMM (cg_gemm_stream.cu) against the pipelined kernel on one box: parity tests, then C4 / C5 / C2.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "gemm or linear or lstm or glstm or stream or partition or c5 or fourier or contract" > $OUT/pytest_stream.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_stream.log
tail -4 $OUT/pytest_stream.log
for c in c4 c5 c2 c3; do
  for on in 1 0; do
    CG_GEMM_STREAM=$on timeout 400 python bench.py --config $c --no-cpu-baseline --no-sweep > $OUT/bench_${c}_stream$on.json 2> $OUT/bench_${c}_stream$on.err
    echo "bench $c stream=$on exit $?"
    python - <<PY
import json
try:
    d=json.load(open('$OUT/bench_${c}_stream$on.json'))
    k=d.get('kernels_ms_per_step',{})
    print('  ms_per_step %.4f value %.1f'%(d['ms_per_step'], d['value']), {n:round(v['ms_per_step'],4) for n,v in k.items() if 'gemm' in n})
except Exception as e:
    print('  no line', e)
PY
  done
done
