#!/bin/bash
# data-parallel all-reduce variants at N GPUs (gpurun --gpus N): args are "ENV=.. --flag .." strings
N=${1:-2}; shift
i=0
for v in "$@"; do i=$((i+1))
  envs=$(echo "$v" | tr ' ' '\n' | grep '=' | tr '\n' ' '); flags=$(echo "$v" | tr ' ' '\n' | grep -v '=' | tr '\n' ' ')
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus $N --no-cpu-baseline --no-sweep $flags > gpurun_out/bench_dp_v$i.json 2> gpurun_out/bench_dp_v$i.err
  python - "$v" gpurun_out/bench_dp_v$i.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], '| gpus', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'], 4), 'e2e', round(d['e2e']['value']))
except Exception as e:
    print(sys.argv[1], 'failed', e); print(open(sys.argv[2].replace('.json', '.err')).read()[-1200:])
PY
done
