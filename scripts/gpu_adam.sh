#!/bin/bash
# native Adam: parity tests and the C4 step
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "adam or lstm or glstm or graphed or gconv or model or fixtures" > $OUT/pytest_adam.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_adam.log
tail -3 $OUT/pytest_adam.log
timeout 400 python bench.py --config c4 --no-cpu-baseline --no-sweep > $OUT/bench_c4_adam.json 2> $OUT/bench_c4_adam.err
python - <<PY
import json
d=json.load(open('$OUT/bench_c4_adam.json'))
k=d.get('kernels_ms_per_step',{})
print('c4 ms_per_step %.4f e2e %.1f launches/step %s'%(d['ms_per_step'], d['e2e']['value'], d.get('native_launches_per_step')), {n:round(v['ms_per_step'],4) for n,v in k.items()})
PY
