#!/bin/bash
# diagnosis of the data-parallel C2 step at N GPUs: per-step device times with and without the L2 flush
N=${1:-8}
i=0
for envs in "CG_BENCH_STEPTIMES=1" "CG_BENCH_STEPTIMES=1 CG_BENCH_NOFLUSH=1"; do i=$((i+1))
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$i bench.py --gpus $N --no-cpu-baseline --no-sweep > gpurun_out/bench_dp8_v$i.json 2> gpurun_out/bench_dp8_v$i.err
  echo "== $envs"; grep "step ms" gpurun_out/bench_dp8_v$i.err | sort | head -3
  python -c "
import json; d=json.loads(open('gpurun_out/bench_dp8_v$i.json').read().strip().splitlines()[-1]); print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4))"
done
