#!/bin/bash
# C5 row partition at N GPUs (gpurun --gpus N): peer-memory halo exchange vs the all-to-all collective
N=${1:-2}; TAG=${2:-r2}
i=0
for ex in ${3:-peer collective}; do i=$((i+1))
  CG_C5_EXCHANGE=$ex timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$i \
     bench.py --gpus $N --config c5 --no-cpu-baseline --steps 10 > gpurun_out/bench_c5_g${N}_${ex}_$TAG.json 2> gpurun_out/bench_c5_g${N}_${ex}_$TAG.err
  python - $ex gpurun_out/bench_c5_g${N}_${ex}_$TAG <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2] + '.json').read().strip().splitlines()[-1])
    print(sys.argv[1], 'gpus', d['n_gpus'], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1),
          {k: (round(v['ms_per_step'], 3), v['launches_per_step']) for k, v in d.get('kernels_ms_per_step', {}).items()},
          'adj', d['config']['adjoint_rel_err'], d['config'].get('halo_rows_rank0'))
except Exception as e:
    print(sys.argv[1], 'failed:', e); print(open(sys.argv[2] + '.err').read()[-2500:])
PY
done
