#!/bin/bash
mkdir -p gpurun_out/rac; O=gpurun_out/rac
for d in 1 2 3; do YB_TS_DBG=$d python tools/dbg_native_determinism.py v8 64 96 2>&1 | grep -E "python vs native|model.12.cv|model.9.cv1.conv|model.15.cv1" | sed "s/^/dbg=$d /"; done > $O/det.txt
cat $O/det.txt
