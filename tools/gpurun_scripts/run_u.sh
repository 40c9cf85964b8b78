#!/bin/bash
mkdir -p gpurun_out/ru; O=gpurun_out/ru
timeout 900 python -m pytest tests/test_trainer_native.py -m gpu -q --no-header -p no:cacheprovider -s -x > $O/native.log 2>&1; echo "rc=$?" >> $O/native.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fp16_pinned.py tests/test_train_step.py -m gpu -q --no-header -p no:cacheprovider -k "v11 or attention" > $O/v11_tests.log 2>&1; echo "rc=$?" >> $O/v11_tests.log
python tools/profile_ops.py v11s 32 > $O/ops_v11s.txt 2>&1
grep -E "worst|passed|failed|^E  |FAILED|rc=|Error" $O/native.log | head -20; tail -2 $O/v11_tests.log; grep -E " other |# layer" $O/ops_v11s.txt
