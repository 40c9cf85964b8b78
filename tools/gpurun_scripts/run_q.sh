#!/bin/bash
mkdir -p gpurun_out/rq; O=gpurun_out/rq
( time python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -s ) > $O/suite.log 2>&1; echo "rc=$?" >> $O/suite.log
python tools/profile_ops.py v11s 32 > $O/ops_v11s.txt 2>&1
python tools/exp_train_profile.py v11s 16 tc > $O/train_profile_tc.txt 2>&1
python bench.py --mode train --steps 5 --warmup 3 > $O/bench_train_tc.json 2> $O/bench_train_tc.err
python bench.py --model v11s --batch 32 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_v11s.json 2> $O/bench_v11s.err
grep -E "head outputs|passed|failed|^E  |FAILED|rc=" $O/suite.log | head; grep -E " other |# layer" $O/ops_v11s.txt; head -24 $O/train_profile_tc.txt | grep -v -i warn; head -c 300 $O/bench_train_tc.json; echo; python -c "
import json; d=json.load(open('$O/bench_v11s.json')); print('v11s', d['value'], d['e2e']['value'])"
