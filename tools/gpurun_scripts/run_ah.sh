#!/bin/bash
mkdir -p gpurun_out/rah; python tools/dbg_u8.py > gpurun_out/rah/u8.txt 2>&1; tail -8 gpurun_out/rah/u8.txt
