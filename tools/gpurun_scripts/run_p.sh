#!/bin/bash
mkdir -p gpurun_out/rp; O=gpurun_out/rp
for a in "v8 320 320 4" "v11 320 320 4" "v8 64 96 2" "v11 64 64 2"; do python tools/exp_torch_tf32.py $a 2>&1 | grep -E "^#|Error" ; done > $O/torch_tf32.txt
timeout 900 python -m pytest tests/test_heads.py tests/test_train_step.py tests/test_gpu_parity.py tests/test_gpu_fp16_pinned.py -m gpu -q --no-header -p no:cacheprovider -s -k "heads or probiou or obb or train_step or v11 or attention" > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
python tools/profile_ops.py v11s 32 > $O/ops_v11s.txt 2>&1
python tools/exp_train_profile.py v11s 16 tc > $O/train_profile_tc.txt 2>&1
cat $O/torch_tf32.txt; grep -E "head outputs|passed|failed|^E  |FAILED" $O/tests.log | head; grep -E " other |# layer" $O/ops_v11s.txt; grep -E "attn|attention|# " $O/train_profile_tc.txt
