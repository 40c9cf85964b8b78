#!/bin/bash
mkdir -p gpurun_out/ras; O=gpurun_out/ras
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_trainer_native.py tests/test_train_step.py -m gpu -q --no-header -p no:cacheprovider -k "bn or native or train" > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
python bench.py --mode train --steps 12 --warmup 3 > $O/bench_train.json 2> $O/bench_train.err
python tools/exp_train_profile.py v11s 16 native > $O/train_profile_native.txt 2>&1
grep -E "passed|failed|^E  |FAILED|rc=" $O/tests.log | head -12; head -c 200 $O/bench_train.json; echo; grep -v -i warn $O/train_profile_native.txt | head -12
