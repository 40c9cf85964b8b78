#!/bin/bash
mkdir -p gpurun_out/ri; O=gpurun_out/ri
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -m gpu -q --no-header -p no:cacheprovider -s > $O/conv_tc.log 2>&1; echo "rc=$?" >> $O/conv_tc.log
grep -E "forward|passed|failed|Error|error|rc=" $O/conv_tc.log | head -60
