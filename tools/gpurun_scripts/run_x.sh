#!/bin/bash
mkdir -p gpurun_out/rx; O=gpurun_out/rx
timeout 900 python -m pytest tests/test_trainer_native.py tests/test_train_step.py tests/test_gpu_conv_tc.py -m gpu -q --no-header -p no:cacheprovider -s -x > $O/native.log 2>&1; echo "rc=$?" >> $O/native.log
YB_NO_SAMPLER=1 YB_STEP_TIMES=1 python bench.py --mode train --steps 12 --warmup 3 > $O/b.json 2> $O/b.err
YB_STEP_TIMES=1 python bench.py --mode train --steps 12 --warmup 3 > $O/a.json 2> $O/a.err
python bench.py --mode train --steps 12 --warmup 3 > $O/c.json 2> $O/c.err
python tools/exp_train_profile.py v11s 16 native > $O/train_profile_native.txt 2>&1
grep -E "worst|passed|failed|^E  |FAILED|rc=|Error" $O/native.log | head -12; grep "step " $O/b.err | tr '\n' ' '; echo; grep "step " $O/a.err | tr '\n' ' '; echo; for f in a b c; do head -c 180 $O/$f.json; echo; done; head -24 $O/train_profile_native.txt | grep -v -i warn
