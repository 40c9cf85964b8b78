#!/bin/bash
mkdir -p gpurun_out/rv; O=gpurun_out/rv
python bench.py --mode train --steps 8 --warmup 3 > $O/bench_train_native.json 2> $O/bench_train_native.err
python bench.py --mode train --train-impl python --steps 5 --warmup 3 > $O/bench_train_python.json 2> $O/bench_train_python.err
python tools/exp_train_profile.py v11s 16 native > $O/train_profile_native.txt 2>&1
head -c 400 $O/bench_train_native.json; echo; tail -2 $O/bench_train_native.err; head -c 300 $O/bench_train_python.json; echo; head -36 $O/train_profile_native.txt | grep -v -i warn
