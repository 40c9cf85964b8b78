#!/bin/bash
mkdir -p gpurun_out/rae; O=gpurun_out/rae
( time python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -s ) > $O/suite.log 2>&1; echo "rc=$?" >> $O/suite.log
python bench.py --mode train --steps 12 --warmup 3 > $O/bench_train.json 2> $O/bench_train.err
python tools/exp_train_profile.py v11s 16 native > $O/train_profile_native.txt 2>&1
python bench.py --model v11s --batch 32 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_v11s.json 2> $O/bench_v11s.err
python bench.py --model v11n --batch 32 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_v11n.json 2> $O/bench_v11n.err
grep -E "worst parameter gradients|passed|failed|^E  |FAILED|rc=" $O/suite.log | head -12; head -c 200 $O/bench_train.json; echo; head -26 $O/train_profile_native.txt | grep -v -i warn
python -c "
import json
for m in ('v11s','v11n'):
    d=json.load(open('$O/bench_%s.json'%m)); print(m, d['value'], d['e2e']['value'])"
