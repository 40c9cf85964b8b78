#!/bin/bash
mkdir -p gpurun_out/rl; O=gpurun_out/rl
python tools/dbg_train_tc.py v8 64 96 2 > $O/dbg_v8_small.txt 2>&1
python tools/dbg_train_tc.py v8 320 320 4 > $O/dbg_v8_big.txt 2>&1
python tools/dbg_train_tc.py v11 64 64 2 > $O/dbg_v11_small.txt 2>&1
cat $O/dbg_v8_small.txt; tail -12 $O/dbg_v8_big.txt; tail -5 $O/dbg_v11_small.txt
