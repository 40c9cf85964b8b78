#!/bin/bash
mkdir -p gpurun_out/rj; O=gpurun_out/rj
timeout 900 python -m pytest tests/test_train_step.py -m gpu -q --no-header -p no:cacheprovider -s > $O/train_tests.log 2>&1; echo "rc=$?" >> $O/train_tests.log
python bench.py --mode train --steps 5 --warmup 3 > $O/bench_train_tc.json 2> $O/bench_train_tc.err
python tools/exp_train_profile.py v11s 16 tc > $O/train_profile_tc.txt 2>&1
grep -E "worst|passed|failed|rc=" $O/train_tests.log | head; cat $O/bench_train_tc.json | head -c 600; echo; tail -2 $O/bench_train_tc.err; head -32 $O/train_profile_tc.txt
