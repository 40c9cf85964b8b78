#!/bin/bash
mkdir -p gpurun_out/rab; O=gpurun_out/rab
python tools/dbg_native_determinism.py v8 64 96 > $O/det.txt 2>&1
grep -A80 "per tensor" $O/det.txt | head -90
