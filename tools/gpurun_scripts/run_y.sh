#!/bin/bash
mkdir -p gpurun_out/ry; O=gpurun_out/ry
timeout 900 python -m pytest tests/test_trainer_native.py -m gpu -q --no-header -p no:cacheprovider -s > $O/native.log 2>&1; echo "rc=$?" >> $O/native.log
grep -E "worst|passed|failed|^E  |FAILED|rc=|Error" $O/native.log | head -12
