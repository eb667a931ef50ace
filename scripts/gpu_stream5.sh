#!/bin/bash
# raw-ring depth sweep of the streaming GEMM at the C5 contraction shape
OUT=gpurun_out
for r in 4 5 6 7; do
  CG_GEMM_STREAM_RAW=$r timeout 400 python bench.py --config c5 --no-cpu-baseline --no-sweep > $OUT/bench_c5_raw$r.json 2> $OUT/bench_c5_raw$r.err
  python - <<PY
import json
d=json.load(open('$OUT/bench_c5_raw$r.json'))
k=d.get('kernels_ms_per_step',{})
print('c5 raw=$r ms_per_step %.4f'%d['ms_per_step'], {n:round(v['ms_per_step'],4) for n,v in k.items() if 'gemm' in n or 'spmm' in n})
PY
done
