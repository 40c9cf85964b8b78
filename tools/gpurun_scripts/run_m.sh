#!/bin/bash
mkdir -p gpurun_out/rm; O=gpurun_out/rm
( time python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -s ) > $O/suite.log 2>&1; echo "rc=$?" >> $O/suite.log
python bench.py --model v11s --batch 32 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_v11s.json 2> $O/bench_v11s.err
python bench.py --model v11n --batch 32 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_v11n.json 2> $O/bench_v11n.err
python tools/profile_ops.py v11s 32 > $O/ops_v11s.txt 2>&1
grep -E "head outputs|passed|failed|rc=|^E  |^FAILED" $O/suite.log | head -20
python -c "
import json
for m in ('v11s','v11n'):
    d=json.load(open('$O/bench_%s.json'%m)); print(m, d['value'], d['e2e']['value'], d['roofline']['frac'])"
grep -E " dw | other |# layer" $O/ops_v11s.txt
