#!/bin/bash
# K tails in the streaming GEMM (parity + C3 A/B) and one ncu --set full capture of k_gemm_stream at the C5 contraction shape
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "gemm or linear or lstm or c3 or newsgroups or sparse" > $OUT/pytest_stream3.log 2>&1
echo "pytest exit $?" >> $OUT/pytest_stream3.log
tail -3 $OUT/pytest_stream3.log
for on in 1 0; do
  CG_GEMM_STREAM=$on timeout 400 python bench.py --config c3 --no-cpu-baseline --no-sweep > $OUT/bench_c3_ktail$on.json 2> $OUT/bench_c3_ktail$on.err
  python - <<PY
import json
d=json.load(open('$OUT/bench_c3_ktail$on.json'))
k=d.get('kernels_ms_per_step',{})
print('c3 stream=$on  ms_per_step %.4f'%d['ms_per_step'], {n:round(v['ms_per_step'],4) for n,v in k.items() if 'gemm' in n})
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_gemm_stream' -s 3 -c 3 \
    -o $OUT/ncu_gemm_stream_c5 -f python bench.py --config c5 --no-cpu-baseline --steps 1 --warmup 1 > $OUT/ncu_gemm_stream_c5.log 2>&1
tail -1 $OUT/ncu_gemm_stream_c5.log
