#!/bin/bash
# ncu captures: tiled attention (inference), TF32 conv / wgrad kernels (training), launch list of one bench step
mkdir -p gpurun_out/rs; O=gpurun_out/rs
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tiled -c 1 -o $O/attn python tools/ncu_target.py v11s 32 > $O/attn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tf_wgrad_kernel -s 100 -c 4 -o $O/wgrad python tools/ncu_train_target.py 16 > $O/wgrad.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tf_conv_kernel -s 200 -c 4 -o $O/tfconv python tools/ncu_train_target.py 16 > $O/tfconv.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_v8n.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-real-weights > $O/bench_under_ncu.log 2>&1
ls -la $O; tail -3 $O/attn.log $O/wgrad.log $O/tfconv.log
