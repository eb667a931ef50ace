#!/bin/bash
# Quick GPU-box check: a pytest selection, then one bench line with per-kernel times.
# Usage (through gpurun): scripts/gpu_quick.sh <tag> "<pytest -k expression or empty>" [bench args...]
TAG=$1; KEXPR=$2; shift 2
mkdir -p gpurun_out
if [ -n "$KEXPR" ]; then
  timeout 900 python -m pytest tests -x -q -m gpu -k "$KEXPR" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_$TAG.log
fi
timeout 600 python bench.py --no-cpu-baseline "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
    print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d.get('e2e', {}).get('value'))
    print({k: round(v['ms_per_step'], 4) for k, v in d.get('kernels_ms_per_step', {}).items()})
    if 'batch_sweep' in d: print({k: round(v['samples_per_s']) for k, v in d['batch_sweep'].items()})
except Exception as e:
    print('no bench line:', e); print(open("gpurun_out/bench_$TAG.err").read()[-2000:])
PY
