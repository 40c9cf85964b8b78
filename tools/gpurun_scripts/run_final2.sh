#!/bin/bash
# end-of-round evidence on the final tree: whole GPU suite, smoke(), default bench line (with roofline.traffic), train bench line
mkdir -p gpurun_out/rf2; O=gpurun_out/rf2
timeout 1500 python -m pytest tests/ -m gpu -q --no-header -p no:cacheprovider > $O/gpu_suite.txt 2>&1; echo "rc=$?" >> $O/gpu_suite.txt
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
python bench.py > $O/bench_default.json 2> $O/bench_default.err
python bench.py --mode train --steps 12 --warmup 3 > $O/bench_train.json 2> $O/bench_train.err
python tools/exp_train_profile.py v11s 16 native > $O/train_profile_native.txt 2>&1
tail -3 $O/gpu_suite.txt; tail -3 $O/smoke.log; head -c 150 $O/bench_default.json; echo; head -c 200 $O/bench_train.json; echo
