#!/bin/bash
mkdir -p gpurun_out/rd; O=gpurun_out/rd
python -m pytest tests/test_gpu_fp16_pinned.py -m gpu -q -s --no-header -p no:cacheprovider > $O/pinned.log 2>&1; echo "rc=$?" >> $O/pinned.log
python -m pytest tests/test_train_step.py -m gpu -q --no-header -p no:cacheprovider > $O/train_tests.log 2>&1; echo "rc=$?" >> $O/train_tests.log
for v in 0 2 1; do
  YB_CHAIN=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-real-weights > $O/bench_chain$v.json 2> $O/bench_chain$v.err
done
YB_CHAIN=1 YB_CHAIN_GRID1=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-real-weights > $O/bench_chain1_g2.json 2> $O/bench_chain1_g2.err
timeout 600 python bench.py --mode train --steps 3 --warmup 1 > $O/bench_train.json 2> $O/bench_train.err
timeout 600 python bench.py --mode train --model v11n --steps 3 --warmup 1 > $O/bench_train_v11n.json 2> $O/bench_train_v11n.err
tail -n 4 $O/pinned.log; tail -n 6 $O/train_tests.log; for f in $O/bench_chain*.json; do echo $f $(python -c "import json,sys; d=json.load(open('$f')); print(d['value'], d['roofline']['forward_ms_per_step'], d['e2e']['value'], d['roofline']['frac'])" 2>&1 | tail -1); done; cat $O/bench_train.json | cut -c1-400; tail -n 3 $O/bench_train.err
