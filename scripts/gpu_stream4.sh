#!/bin/bash
# deeper raw ring of the streaming GEMM: parity, then C5 / C4 / C2
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "gemm or linear or lstm" > $OUT/pytest_stream4.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_stream4.log
tail -3 $OUT/pytest_stream4.log
for c in c5 c4 c2; do
  timeout 400 python bench.py --config $c --no-cpu-baseline --no-sweep > $OUT/bench_${c}_raw8.json 2> $OUT/bench_${c}_raw8.err
  python - <<PY
import json
d=json.load(open('$OUT/bench_${c}_raw8.json'))
k=d.get('kernels_ms_per_step',{})
print('$c ms_per_step %.4f'%d['ms_per_step'], {n:round(v['ms_per_step'],4) for n,v in k.items() if 'gemm' in n or 'spmm' in n})
PY
done
