#!/bin/bash
mkdir -p gpurun_out/rad; O=gpurun_out/rad
for d in 0 1 2 3; do YB_TS_UP=$d python tools/dbg_native_determinism.py v8 64 96 2>&1 | grep -E "python vs native|model.12.cv2.conv|model.9.cv1.conv|model.2.cv1.conv" | sed "s/^/up=$d /"; done > $O/det.txt
cat $O/det.txt
