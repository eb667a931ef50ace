#!/bin/bash
# Multi-GPU pass (run with gpurun --gpus N): C2 weak-scaling line and C5 row-partitioned line at N GPUs
N=${1:-2}; TAG=${2:-r2}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
timeout 600 bash -c "$(declare -f run); N=$N; run --no-cpu-baseline --no-sweep" > gpurun_out/bench_c2_g${N}_$TAG.json 2> gpurun_out/bench_c2_g${N}_$TAG.err; echo "c2 x$N exit $?"
timeout 600 bash -c "$(declare -f run); N=$N; run --config c5 --no-cpu-baseline --steps 5" > gpurun_out/bench_c5_g${N}_$TAG.json 2> gpurun_out/bench_c5_g${N}_$TAG.err; echo "c5 x$N exit $?"
python - <<PY
import json
for c in ('c2', 'c5'):
    try:
        d = json.loads(open('gpurun_out/bench_%s_g${N}_$TAG.json' % c).read().strip().splitlines()[-1])
        print(c, 'gpus', d['n_gpus'], 'value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1),
              {k: (round(v['ms_per_step'], 3), v['launches_per_step']) for k, v in d.get('kernels_ms_per_step', {}).items() if v['ms_per_step'] > 0.05},
              d['config'].get('exchange'), d['config'].get('halo_rows_rank0'))
    except Exception as e:
        print(c, 'failed:', e); print(open('gpurun_out/bench_%s_g${N}_$TAG.err' % c).read()[-1500:])
PY
