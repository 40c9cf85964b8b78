#!/bin/bash
mkdir -p gpurun_out/raq; O=gpurun_out/raq
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_trainer_native.py tests/test_train_step.py tests/test_gpu_conv_tc.py -m gpu -q --no-header -p no:cacheprovider -k "bn or native or train or conv_tc or stem or wgrad or backward" > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
python bench.py --mode train --steps 12 --warmup 3 > $O/bench_train.json 2> $O/bench_train.err
YB_NO_PDL=1 python bench.py --mode train --steps 12 --warmup 3 > $O/bench_train_nopdl.json 2> $O/bench_train_nopdl.err
python tools/dbg_native_determinism.py > $O/determinism.txt 2>&1
grep -E "passed|failed|^E  |FAILED|rc=" $O/tests.log | head -12; head -c 200 $O/bench_train.json; echo; head -c 200 $O/bench_train_nopdl.json; echo; tail -5 $O/determinism.txt
