#!/bin/bash
mkdir -p gpurun_out/rap; O=gpurun_out/rap
timeout 600 python -m pytest tests/test_metrics.py -q --no-header -p no:cacheprovider > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
grep -E "AssertionError|passed|failed" $O/tests.log | cut -c1-700 | tail -20; tail -3 $O/tests.log
