#!/bin/bash
# C5 streaming-step variants side by side: bench --config c5 per environment setting (args: "VAR=val VAR=val" ...)
if [ -n "${RUN_TESTS:-}" ]; then timeout 600 python -m pytest tests -x -q -m gpu -k "streaming or basis or partitioned or filter or row_block or c3 or c1" 2>&1 | tail -3; fi
i=0
for envs in "$@"; do i=$((i+1))
  env $envs timeout 600 python bench.py --config c5 --no-cpu-baseline --steps 5 > gpurun_out/bench_c5_v$i.json 2> gpurun_out/bench_c5_v$i.err
  python - "$envs" gpurun_out/bench_c5_v$i.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d["ms_per_step"], 3), {k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels_ms_per_step"].items()},
      round(d["roofline"]["frac"], 3), d["config"]["adjoint_rel_err"])
PY
done
