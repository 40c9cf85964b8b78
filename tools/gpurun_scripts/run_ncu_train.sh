#!/bin/bash
# ncu --set full of the big-layer launches of the training step's memory-bound kernels and of the wgrad kernel (final state)
mkdir -p gpurun_out/rnt; O=gpurun_out/rnt
N="ncu --set full --clock-control none --import-source on"
timeout 400 $N -k regex:bn_stats4 -s 162 -c 2 -o $O/bn_stats_fwd python tools/ncu_train_target.py 16 > $O/a.log 2>&1
timeout 400 $N -k regex:bn_stats4 -s 322 -c 2 -o $O/bn_stats_bwd python tools/ncu_train_target.py 16 > $O/b.log 2>&1
timeout 400 $N -k regex:tf_wgrad_kernel -s 155 -c 3 -o $O/wgrad_final python tools/ncu_train_target.py 16 > $O/c.log 2>&1
timeout 400 $N -k regex:bn_silu_dz4 -s 160 -c 2 -o $O/bn_dz python tools/ncu_train_target.py 16 > $O/d.log 2>&1
ls -la $O | head; tail -2 $O/a.log
