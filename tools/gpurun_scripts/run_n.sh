#!/bin/bash
mkdir -p gpurun_out/rn; O=gpurun_out/rn
python tools/dbg_train_tc.py v8 64 96 2 > $O/dbg_v8_small.txt 2>&1
python tools/dbg_train_tc.py v8 320 320 4 > $O/dbg_v8_big.txt 2>&1
python tools/dbg_train_tc.py v11 64 64 2 > $O/dbg_v11_small.txt 2>&1
python tools/dbg_train_tc.py v11 320 320 4 > $O/dbg_v11_big.txt 2>&1
grep "flat gradient" $O/*.txt; grep -A70 "per conv weight" $O/dbg_v8_big.txt | head -75
