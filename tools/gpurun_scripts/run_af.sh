#!/bin/bash
# ncu: the big (early-layer) launches of the TF32 conv kernels in the native training step
mkdir -p gpurun_out/raf; O=gpurun_out/raf
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tf_conv_kernel -s 177 -c 8 -o $O/tfconv_big python tools/ncu_train_target.py 16 > $O/tfconv.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tf_wgrad_kernel -s 150 -c 10 -o $O/wgrad_big python tools/ncu_train_target.py 16 > $O/wgrad.log 2>&1
ls -la $O; tail -n 3 $O/tfconv.log; tail -n 3 $O/wgrad.log
