#!/bin/bash
# 8-GPU line of the training step (native step, NCCL all-reduce of the flat gradient buffer between backward and apply)
mkdir -p gpurun_out/r8; O=gpurun_out/r8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --mode train --steps 12 --warmup 3 > $O/train8.json 2> $O/train8.err
grep -h "metric" $O/train8.json | head -c 700; echo; tail -3 $O/train8.err
