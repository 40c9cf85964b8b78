#!/bin/bash
mkdir -p gpurun_out/rao; O=gpurun_out/rao
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_trainer_native.py -m gpu -q --no-header -p no:cacheprovider > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
python bench.py --mode train --steps 12 --warmup 3 > $O/bench_train.json 2> $O/bench_train.err
YB_WGRAD_NO_MERGE=1 python bench.py --mode train --steps 12 --warmup 3 > $O/bench_train_nomerge.json 2> $O/bench_train_nomerge.err
YB_PROF_DUMP=tf_wgrad_kernel,tf_conv_kernel python tools/exp_train_profile.py v11s 16 native > $O/train_profile_native.txt 2>&1
grep -E "passed|failed|^E  |FAILED|rc=" $O/tests.log | head -12; head -c 200 $O/bench_train.json; echo; head -c 200 $O/bench_train_nomerge.json; echo; grep -v -i warn $O/train_profile_native.txt | head -6
