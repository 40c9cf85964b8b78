#!/bin/bash
mkdir -p gpurun_out/rl; O=gpurun_out/rl
timeout 60 python -m pytest tests/test_gpu_parity.py tests/test_trainer_native.py -m gpu -q --no-header -p no:cacheprovider -k "loss or native_train_step_matches" > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
tail -3 $O/tests.log
