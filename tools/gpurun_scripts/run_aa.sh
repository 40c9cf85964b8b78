#!/bin/bash
mkdir -p gpurun_out/raa; O=gpurun_out/raa
python tools/dbg_native_determinism.py v8 128 160 > $O/det.txt 2>&1
python tools/dbg_native_determinism.py v11 128 128 >> $O/det.txt 2>&1
python tools/dbg_native_determinism.py v8 64 96 >> $O/det.txt 2>&1
grep -E "python|native|Error" $O/det.txt
