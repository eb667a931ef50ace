#!/bin/bash
OUT=gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for i in 1 2; do
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_final$i.log 2>&1; echo "full run $i exit $?"; tail -2 $OUT/pytest_final$i.log; grep -c "AccumulateGrad node's stream does not match" $OUT/pytest_final$i.log
done
