#!/bin/bash
mkdir -p gpurun_out/rr; O=gpurun_out/rr
timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_train_step.py -m gpu -q --no-header -p no:cacheprovider -s > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
python tools/exp_train_profile.py v11s 16 tc > $O/train_profile_tc.txt 2>&1
python bench.py --mode train --steps 5 --warmup 3 > $O/bench_train_tc.json 2> $O/bench_train_tc.err
grep -E "forward [0-9]|head outputs|passed|failed|^E  |FAILED|rc=" $O/tests.log | head -24; head -20 $O/train_profile_tc.txt | grep -v -i warn; head -c 300 $O/bench_train_tc.json; echo
