#!/bin/bash
mkdir -p gpurun_out/rag; O=gpurun_out/rag
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_trainer_native.py tests/test_train_step.py -m gpu -q --no-header -p no:cacheprovider -s > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
python bench.py --mode train --steps 12 --warmup 3 > $O/bench_train.json 2> $O/bench_train.err
YB_TF_OCC1=1 python bench.py --mode train --steps 12 --warmup 3 > $O/bench_train_occ1.json 2> $O/bench_train_occ1.err
python tools/exp_train_profile.py v11s 16 native > $O/train_profile_native.txt 2>&1
grep -E "worst parameter gradients|head outputs|passed|failed|^E  |FAILED|rc=" $O/tests.log | head -12; head -c 200 $O/bench_train.json; echo; head -c 200 $O/bench_train_occ1.json; echo; head -16 $O/train_profile_native.txt | grep -v -i warn
