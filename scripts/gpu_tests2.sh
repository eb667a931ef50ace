#!/bin/bash
OUT=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "deferred" > $OUT/pytest_defA.log 2>&1; echo "alone exit $?"; tail -2 $OUT/pytest_defA.log
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_fullB.log 2>&1; echo "full exit $?"; tail -5 $OUT/pytest_fullB.log
