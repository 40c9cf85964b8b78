#!/bin/bash
mkdir -p gpurun_out/ro; O=gpurun_out/ro
python tools/exp_torch_tf32.py v8 320 320 4 > $O/torch_tf32.txt 2>&1
python tools/exp_torch_tf32.py v11 320 320 4 >> $O/torch_tf32.txt 2>&1
python tools/exp_torch_tf32.py v8 64 96 2 >> $O/torch_tf32.txt 2>&1
python tools/exp_torch_tf32.py v11 64 64 2 >> $O/torch_tf32.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fp16_pinned.py tests/test_train_step.py -m gpu -q --no-header -p no:cacheprovider -k "v11 or attention" > $O/v11_tests.log 2>&1; echo "rc=$?" >> $O/v11_tests.log
python tools/profile_ops.py v11s 32 > $O/ops_v11s.txt 2>&1
python tools/exp_train_profile.py v11s 16 tc > $O/train_profile_tc.txt 2>&1
grep "#" $O/torch_tf32.txt | tail -8; tail -3 $O/v11_tests.log; grep -E " other |# layer|# v11s" $O/ops_v11s.txt; head -12 $O/train_profile_tc.txt | grep -v Warn
