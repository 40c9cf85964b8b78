#!/bin/bash
mkdir -p gpurun_out/rk; O=gpurun_out/rk
timeout 900 python -m pytest tests/test_train_step.py tests/test_topk.py tests/test_gpu_conv_tc.py -m gpu -q --no-header -p no:cacheprovider -s > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
python bench.py --mode train --steps 5 --warmup 3 > $O/bench_train_tc.json 2> $O/bench_train_tc.err
python tools/exp_train_profile.py v11s 16 tc > $O/train_profile_tc.txt 2>&1
python tools/profile_ops.py v11s 32 > $O/ops_v11s.txt 2>&1
grep -E "worst|head outputs|passed|failed|rc=|^E  " $O/tests.log | head -20; cat $O/bench_train_tc.json | head -c 400; echo; tail -2 $O/bench_train_tc.err; head -30 $O/train_profile_tc.txt
