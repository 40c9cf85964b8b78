#!/bin/bash
mkdir -p gpurun_out/rsg; O=gpurun_out/rsg
timeout 300 python -m pytest tests/test_segloss.py -q --no-header -p no:cacheprovider > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
grep -E "Error|passed|failed|assert" $O/tests.log | cut -c1-400 | tail -12
