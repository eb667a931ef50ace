#!/bin/bash
# fused-kernel phase stamps for both gather forms (CTA 0, second group)
for b in 0 1; do
  echo "== CG_FUSED_BLOCK=$b"
  CG_FUSED_BLOCK=$b CG_TRACE=1 timeout 300 python scripts/prof_fused.py --iters 3 --kernels 1 "$@" 2>&1 | tail -${TAILN:-40}
done
