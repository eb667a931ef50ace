#!/bin/bash
# C5 streaming-step forms side by side: tests of the streaming paths, then bench --config c5 per form
timeout 600 python -m pytest tests -x -q -m gpu -k "streaming or basis or partitioned or filter_forward_backward" 2>&1 | tail -4
for mode in "0 0" "1 0" "1 1"; do set -- $mode
  CG_SPMM_BLOCK=$1 CG_SPMM_TILE=$2 timeout 600 python bench.py --config c5 --no-cpu-baseline --steps 5 > gpurun_out/bench_c5_m$1$2.json 2> gpurun_out/bench_c5_m$1$2.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_c5_m$1$2.json").read().strip().splitlines()[-1])
print("block $1 tile $2", round(d["ms_per_step"], 3), {k: (round(v["ms_per_step"], 3), v["launches_per_step"]) for k, v in d["kernels_ms_per_step"].items()},
      round(d["roofline"]["frac"], 3), d["config"]["adjoint_rel_err"])
PY
done
