#!/bin/bash
# persistent k_basis_onchip: parity tests and the C4 / C2 / C1 steps
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "basis or filter or lstm or glstm or fixtures or first_layer" > $OUT/pytest_basis.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_basis.log
tail -3 $OUT/pytest_basis.log
for c in c4 c2 c1; do
timeout 400 python bench.py --config $c --no-cpu-baseline --no-sweep > $OUT/bench_${c}_basis.json 2> $OUT/bench_${c}_basis.err
python - <<PY
import json
d=json.load(open('$OUT/bench_${c}_basis.json'))
k=d.get('kernels_ms_per_step',{})
print('$c ms_per_step %.4f'%(d['ms_per_step']), {n:round(v['ms_per_step'],4) for n,v in k.items() if 'basis' in n or 'fused' in n})
PY
done
