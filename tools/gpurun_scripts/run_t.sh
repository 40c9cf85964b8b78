#!/bin/bash
mkdir -p gpurun_out/rt; O=gpurun_out/rt
timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_train_step.py tests/test_heads.py -m gpu -q --no-header -p no:cacheprovider -s > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fp16_pinned.py -m gpu -q --no-header -p no:cacheprovider -k "v11 or attention" > $O/v11_tests.log 2>&1; echo "rc=$?" >> $O/v11_tests.log
python tools/exp_train_profile.py v11s 16 tc > $O/train_profile_tc.txt 2>&1
python bench.py --mode train --steps 5 --warmup 3 > $O/bench_train_tc.json 2> $O/bench_train_tc.err
python tools/profile_ops.py v11s 32 > $O/ops_v11s.txt 2>&1
python bench.py --model v11s --batch 32 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_v11s.json 2> $O/bench_v11s.err
grep -E "wgrad [0-9]|head outputs|passed|failed|^E  |FAILED|rc=" $O/tests.log | head -24; tail -2 $O/v11_tests.log; head -16 $O/train_profile_tc.txt | grep -v -i warn; head -c 300 $O/bench_train_tc.json; echo; grep -E " other |# layer" $O/ops_v11s.txt; python -c "
import json; d=json.load(open('$O/bench_v11s.json')); print('v11s', d['value'], d['e2e']['value'])"
