#!/bin/bash
mkdir -p gpurun_out/rz; O=gpurun_out/rz
timeout 900 python -m pytest tests/test_trainer_native.py tests/test_train_step.py -m gpu -q --no-header -p no:cacheprovider -s > $O/native.log 2>&1; echo "rc=$?" >> $O/native.log
python tools/exp_train_profile.py v11s 16 native > $O/train_profile_native.txt 2>&1
python bench.py --mode train --steps 12 --warmup 3 > $O/bench_train.json 2> $O/bench_train.err
grep -E "worst|passed|failed|^E  |FAILED|rc=|Error" $O/native.log | head -12; head -c 200 $O/bench_train.json; echo; grep -E "attn|# " $O/train_profile_native.txt
