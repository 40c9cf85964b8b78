#!/bin/bash
mkdir -p gpurun_out/rf1; O=gpurun_out/rf1
timeout 1500 python -m pytest tests/ -m gpu -q --no-header -p no:cacheprovider > $O/gpu_suite.txt 2>&1; echo "rc=$?" >> $O/gpu_suite.txt
python bench.py --mode train --steps 12 --warmup 3 > $O/bench_train.json 2> $O/bench_train.err
python tools/dbg_native_determinism.py > $O/determinism.txt 2>&1
python bench.py > $O/bench_default.json 2> $O/bench_default.err
tail -4 $O/gpu_suite.txt; head -c 200 $O/bench_train.json; echo; head -4 $O/determinism.txt; head -c 300 $O/bench_default.json; echo
