#!/bin/bash
# GPU run B (1 GPU): suite with upsample fusion + H3, parity table for tanh / exact SiLU, e2e matrix, H3 per-op sweep
mkdir -p gpurun_out/rb; O=gpurun_out/rb
python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x --deselect tests/test_gpu_fp16_pinned.py > $O/suite.log 2>&1; echo "suite rc=$?" >> $O/suite.log
YB_PRINT_LAYER_TABLE=1 YB_SILU=tanh python -m pytest tests/test_gpu_fp16_pinned.py -m gpu -q -s --no-header -p no:cacheprovider > $O/pinned_tanh.log 2>&1; echo "rc=$?" >> $O/pinned_tanh.log
YB_PRINT_LAYER_TABLE=1 YB_SILU=exact python -m pytest tests/test_gpu_fp16_pinned.py -m gpu -q -s --no-header -p no:cacheprovider > $O/pinned_exact.log 2>&1; echo "rc=$?" >> $O/pinned_exact.log
python tools/exp_e2e.py default > $O/e2e.log 2>&1
OMP_NUM_THREADS=1 python tools/exp_e2e.py omp1 >> $O/e2e.log 2>&1
YB_DEBUG_NO_PRIORITY=1 python tools/exp_e2e.py no_priority >> $O/e2e.log 2>&1
CUDA_DEVICE_MAX_CONNECTIONS=32 python tools/exp_e2e.py maxconn32 >> $O/e2e.log 2>&1
python tools/exp_e2e.py nccl_init --nccl >> $O/e2e.log 2>&1
for cfg in "0 0" "64 0" "32 0" "64 32" "16 0"; do
  set -- $cfg
  for m in v8n:32 v8s:32 v8x:8; do
    YB_H3_MIN_N=$1 YB_H3_BK=$2 python bench.py --model ${m%%:*} --batch ${m##*:} --steps 20 --warmup 5 --no-cpu-baseline --no-real-weights > $O/h3_${1}_${2}_${m%%:*}.json 2>> $O/h3.err
  done
done
YB_SILU=exact python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-real-weights > $O/silu_exact_v8n.json 2>> $O/h3.err
for m in v8n:32 v8x:8; do
  YB_H3_MIN_N=64 python tools/profile_ops.py ${m%%:*} ${m##*:} > $O/ops_h3_${m%%:*}.txt 2>&1
  YB_H3_MIN_N=0 python tools/profile_ops.py ${m%%:*} ${m##*:} > $O/ops_h0_${m%%:*}.txt 2>&1
done
tail -2 $O/suite.log; cat $O/e2e.log; for f in $O/h3_*.json $O/silu_exact_v8n.json; do echo $f $(python -c "import json,sys; d=json.load(open('$f')); print(d['value'], d['roofline']['forward_ms_per_step'])" 2>&1 | tail -1); done
